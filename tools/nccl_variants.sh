#!/bin/bash
# usage: tools/nccl_variants.sh NGPUS "VAR=VAL" ...   -> one line per variant: images/s and ms/step
N=$1; shift
for v in "$@"; do
  out=$(env $v timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>/dev/null | grep '^{' | tail -1)
  echo "$v $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["ms_per_step"],2), "graph" if d["config"]["cuda_graph"] else "eager")
except Exception as e: print("FAILED")')"
done
