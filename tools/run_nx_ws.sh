#!/bin/bash
# step time at N ranks with / without the weight gradients on the side stream (run under `gpurun --gpus N`)
N=${1:-2}
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --parts none --no-cpu-baseline 2>gpurun_out/nx_ws.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=$N $*', 'ms/step', round(d['ms_per_step'],3), 'img/s', round(d['value']), 'identical', d.get('replicas_identical_after_timed_steps'), 'timeouts', d['config'].get('bn_exchange_timeouts'))"
}
run PO2_WGRAD_STREAM=0
run PO2_WGRAD_STREAM=1
