#!/bin/bash
# N=2 step time under variants of the SyncBatchNorm exchange (run under `gpurun --gpus 2`)
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --parts none --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'ms/step', round(d['ms_per_step'],3), 'launches', d['gpu_launches_per_step'])"
}
run PO2_X=0
run PO2_BN_EXCHANGE=local
run PO2_BN_EXCHANGE=nccl
run PO2_BN_FUSED_MULTI=1
echo "N=1 without the one-launch norms:"
PO2_BN_FUSED=0 python bench.py --parts none --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('N=1 PO2_BN_FUSED=0 ms/step', round(d['ms_per_step'],3), d['gpu_launches_per_step'])"
