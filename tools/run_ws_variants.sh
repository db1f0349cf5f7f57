#!/bin/bash
# A/B switches of the N=1 step: tools/run_ws_variants.sh "ENV=.. ENV=.." ...  (one bench run per argument)
run() { name=$1; shift; env "$@" python bench.py --steps 30 --warmup 5 --parts none --no-cpu-baseline > gpurun_out/ws_$name.json 2> gpurun_out/ws_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ws_$name.json").read().strip().splitlines()[-1])
    print("$*", round(d["ms_per_step"], 4), round(d["value"]), d["gpu_launches_per_step"])
except Exception as e:
    print("$*", "failed", e)
PY
}
i=0
for v in "$@"; do i=$((i+1)); run v$i $v; done
