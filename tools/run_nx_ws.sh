#!/bin/bash
# step time at N ranks under a few switches (run under `gpurun --gpus N`): tools/run_nx_ws.sh N "ENV=.. ENV=.." ...
N=${1:-2}; shift
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --parts none --no-cpu-baseline 2>gpurun_out/nx_ws.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=$N $*', 'ms/step', round(d['ms_per_step'],3), 'img/s', round(d['value']), 'identical', d.get('replicas_identical_after_timed_steps'), 'timeouts', d['config'].get('bn_exchange_timeouts'))"
}
if [ $# -eq 0 ]; then run X=1; fi
for v in "$@"; do run $v; done
