#!/bin/bash
# step time at N ranks with / without overlapping the gradient all-reduce with backward (run under `gpurun --gpus N`)
N=${1:-8}
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --parts none --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=$N $*', 'ms/step', round(d['ms_per_step'],3), 'img/s', round(d['value']), 'identical', d.get('replicas_identical_after_timed_steps'))"
}
run PO2_GRAD_OVERLAP=0
run PO2_GRAD_OVERLAP=1
run PO2_GRAD_OVERLAP=1 PO2_GRAD_BUCKETS=8
