#!/bin/bash
# N-rank bench variants: gradient exchange (coalesced all-reduce vs torch DDP) x SyncBatchNorm exchange
N=${1:-2}
shift
for v in "${@:-peer:0 peer:1 nccl:0 local:0}"; do for mv in $v; do
  mode=${mv%%:*}; ddp=${mv##*:}
  PO2_DDP=$ddp PO2_BN_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n${N}_${mode}_ddp$ddp.json 2> gpurun_out/bench_n${N}_${mode}_ddp$ddp.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_${mode}_ddp$ddp.json").read().strip().splitlines()[-1])
    print("bn=$mode ddp=$ddp", round(d["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["value"]), d["config"]["parallelism"])
except Exception as e:
    print("bn=$mode ddp=$ddp FAILED", e)
    print(open("gpurun_out/bench_n${N}_${mode}_ddp$ddp.err").read()[-1500:])
PY
done; done
