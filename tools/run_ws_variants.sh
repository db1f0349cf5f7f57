#!/bin/bash
# A/B of the side-stream weight gradients on the N=1 step: tools/run_ws_variants.sh  (writes gpurun_out/ws_*.json)
run() { name=$1; shift; env "$@" python bench.py --steps 30 --warmup 5 --parts none --no-cpu-baseline > gpurun_out/ws_$name.json 2> gpurun_out/ws_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ws_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"], 4), round(d["value"]), round(d["e2e"]["value"]))
except Exception as e:
    print("$name", "failed", e)
PY
}
run on X=1
run on_bwd2k PO2_BN_FUSED_BWD=0
run on_bn2k PO2_BN_FUSED=0
run off_bn2k PO2_BN_FUSED=0 PO2_WGRAD_STREAM=0
