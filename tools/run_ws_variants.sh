#!/bin/bash
# A/B switches of the N=1 step: tools/run_ws_variants.sh  (writes gpurun_out/ws_*.json)
run() { name=$1; shift; env "$@" python bench.py --steps 30 --warmup 5 --parts none --no-cpu-baseline > gpurun_out/ws_$name.json 2> gpurun_out/ws_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ws_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"], 4), round(d["value"]), round(d["e2e"]["value"]), d["gpu_launches_per_step"])
except Exception as e:
    print("$name", "failed", e)
PY
}
run convbn_off PO2_CONV_BN=0
run convbn_on PO2_CONV_BN=1
