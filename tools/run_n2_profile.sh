#!/bin/bash
# torch.profiler kernel tables of eager steps at N ranks (default bucket-view DDP, and DDP defaults)
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --torch-profile gpurun_out/torch_profile_n$N.txt > gpurun_out/tp_n$N.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/tp_n$N.log
head -40 gpurun_out/torch_profile_n$N.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r01_n${N}_peer.json 2> gpurun_out/bench_r01_n${N}_peer.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r01_n${N}_peer.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
