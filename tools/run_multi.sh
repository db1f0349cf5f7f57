#!/bin/bash
# N-GPU validation (run under `gpurun --gpus N`): the cross-GPU pytest (SyncBN exchange, gradient averaging, sharded
# quantizer on the CUDA kernels), then bench.py at N ranks
N=${1:-2}
TAG=${2:-r02}
python -m pytest tests/test_multigpu_gpu.py -m gpu -q -k "$N" 2>&1 | tail -3
cp gpurun_out/check_cross_gpu_n$N.log gpurun_out/${TAG}_check_cross_gpu_n$N.log 2>/dev/null
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches_per_step"], "replicas_identical", d.get("replicas_identical_after_timed_steps"), "timeouts", d["config"].get("bn_exchange_timeouts"))
PY
grep -i 'error\|Traceback' gpurun_out/${TAG}_bench_n$N.err | head -5
