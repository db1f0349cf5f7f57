#!/bin/bash
# N=2 step time under variants of the gradient all-reduce (run under `gpurun --gpus N`)
N=${1:-2}
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --parts none --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'ms/step', round(d['ms_per_step'],3), 'launches', d['gpu_launches_per_step'])"
}
run PO2_GRAD_BUCKETS=4
run PO2_GRAD_BUCKETS=1
run PO2_GRAD_BUCKETS=2
run PO2_GRAD_BUCKETS=8
run PO2_GRAD_OVERLAP=0
run PO2_GRAD_BUCKETS=1 NCCL_MAX_NCHANNELS=2
run PO2_GRAD_BUCKETS=2 NCCL_MAX_NCHANNELS=4
