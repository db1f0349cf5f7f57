#!/bin/bash
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/check_sync_bn.py > gpurun_out/check_sync_bn_n$N.log 2>&1
echo "check rc=$?"
grep -v '^\*\*\|OMP_NUM' gpurun_out/check_sync_bn_n$N.log | grep "^{" 
