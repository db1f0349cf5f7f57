#!/bin/bash
# 2-GPU validation: SyncBN exchange check, then bench at N=2 (run under `gpurun --gpus 2`)
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/check_sync_bn.py > gpurun_out/check_sync_bn_n$N.log 2>&1
echo "check rc=$?"
grep -v '^\*\*\|OMP_NUM' gpurun_out/check_sync_bn_n$N.log | tail -15
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r01_n${N}_peer.json 2> gpurun_out/bench_r01_n${N}_peer.err
echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r01_n${N}_peer.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["gpu_launches_per_step"], d["config"])
PY
grep -i 'po2\]\|error' gpurun_out/bench_r01_n${N}_peer.err | head
