"""Cross-GPU check of FusedSyncBatchNorm: under torchrun with N ranks, every rank holds a different
shard of one batch; forward output and input gradient must equal single-device batch norm over the
whole batch (computed by every rank in fp64 on the CPU), through both exchange paths.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_sync_bn.py

Prints one JSON line per mode on rank 0 and exits non-zero on any mismatch.  Also replays the
exchange inside a CUDA graph (what bench.py does)."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po2_quantization_b200 as P  # noqa: E402
from po2_quantization_b200 import batchnorm as bnm  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    ok = True
    keep = []                                                      # never free a mailbox while graphs reference it
    for mode in ("peer", "nccl"):
        os.environ["PO2_BN_EXCHANGE"] = mode
        keep.append(dict(bnm._exchanges))
        bnm._exchanges.clear()
        worst = 0.0
        for it, (Bper, C, H, W) in enumerate([(16, 16, 32, 32), (8, 64, 8, 8), (4, 130, 3, 3), (32, 32, 16, 16)]):
            g = torch.Generator().manual_seed(100 + it)
            xa = torch.randn(world * Bper, C, H, W, generator=g) * 1.5 + 0.3
            xa[:Bper] += 2.0                                       # ranks see different distributions
            ra = torch.randn(world * Bper, C, H, W, generator=g)
            ga = torch.randn(world * Bper, C, H, W, generator=g)
            bn = P.FusedSyncBatchNorm(C).cuda().train()
            sl = slice(rank * Bper, (rank + 1) * Bper)
            x = xa[sl].cuda().requires_grad_(True)
            r = ra[sl].cuda().requires_grad_(True)
            y = bn(x, r, True)
            y.backward(ga[sl].cuda())
            xd, rd = xa.double().requires_grad_(True), ra.double().requires_grad_(True)
            yr = F.relu(F.batch_norm(xd, None, None, bn.weight.double().cpu(), bn.bias.double().cpu(), True, 0.0, bn.eps) + rd)
            yr.backward(ga.double())
            for mine, ref in ((y, yr[sl]), (x.grad, xd.grad[sl]), (r.grad, rd.grad[sl])):
                e = ((mine.detach().double().cpu() - ref.detach()).abs().max() / ref.detach().abs().max()).item()
                worst = max(worst, e)
        ex = bnm._exchanges.get(id(dist.group.WORLD))
        keep.append((gr, bn, yy, gx) if False else None)
        used = "peer" if ex is not None else "nccl"
        # graph replay of forward+backward with fresh data each replay
        bn = P.FusedSyncBatchNorm(32).cuda().train()
        xs = torch.randn(8, 32, 16, 16, device="cuda")
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            xi = xs.detach().requires_grad_(True)
            bn(xi, None, True).sum().backward()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                xi = xs.detach().requires_grad_(True)
                yy = bn(xi, None, True)
                gx, = torch.autograd.grad(yy, xi, torch.ones_like(yy))
        torch.cuda.current_stream().wait_stream(s)
        rep_err = 0.0
        for rep in range(6):
            g = torch.Generator().manual_seed(500 + rep)
            full = torch.randn(world * 8, 32, 16, 16, generator=g) * (1 + rep)
            xs.copy_(full[rank * 8:(rank + 1) * 8])
            gr.replay()
            torch.cuda.synchronize()
            ref = F.relu(F.batch_norm(full.double(), None, None, bn.weight.double().cpu(), bn.bias.double().cpu(), True, 0.0, bn.eps))
            rep_err = max(rep_err, ((yy.double().cpu() - ref[rank * 8:(rank + 1) * 8]).abs().max() / ref.abs().max()).item())
        t = torch.tensor([worst, rep_err], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        err_flag = ex.error_flag() if ex is not None else 0
        if rank == 0:
            print(json.dumps({"requested": mode, "used": used, "world": world, "max_rel_err": t[0].item(),
                              "graph_replay_max_rel_err": t[1].item(), "timeout_flag": err_flag}), flush=True)
        ok = ok and t[0].item() < 5e-5 and t[1].item() < 5e-5 and err_flag == 0 and (mode == "nccl" or True)
    # BatchSharded: bucketed, overlapped gradient averaging == plain all_reduce(SUM) / world
    from po2_quantization_b200.distributed import BatchSharded
    torch.manual_seed(7 + rank)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 8, 3, padding=1),
                              torch.nn.Flatten(), torch.nn.Linear(8 * 8 * 8, 10)).cuda()
    import copy
    ref = copy.deepcopy(net)                              # un-wrapped replica: its gradients stay local
    model = BatchSharded(net, overlap=True, buckets=3)    # the overlapped form is the harder one to get right
    ref.load_state_dict(net.state_dict())                 # rank 0's parameters, as broadcast by the wrapper
    xb = torch.randn(4, 3, 8, 8, device="cuda")
    worst = 0.0
    for it in range(3):
        net.zero_grad()
        ref.zero_grad()
        model(xb * (it + 1)).square().mean().backward()   # all-reduces start from the hooks, under backward
        nred = model.average_gradients()
        ref(xb * (it + 1)).square().mean().backward()
        local = [p.grad.clone() for p in ref.parameters()]
        torch.cuda.synchronize()
        for g in local:
            dist.all_reduce(g)
        for p, g in zip(net.parameters(), local):
            worst = max(worst, ((p.grad - g / world).abs().max() / (g.abs().max() / world + 1e-30)).item())
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"batch_sharded_grad_avg_max_rel_err": t.item(), "tensors": nred, "buckets": len(model._buckets)}), flush=True)
    ok = ok and t.item() < 1e-6
    # sharded_quantize on the sm_100a kernels: every rank quantizes its shard of
    # ONE tensor after a single all_reduce(MAX) of the scale; the result must equal the single-device
    # quantizer on the whole tensor, bit for bit -- also when a NaN sits in another rank's shard
    from po2_quantization_b200.distributed import sharded_quantize
    bad = 0
    for case, (n, plus, bits, dt) in enumerate([(1 << 20, False, 4, torch.float32), ((1 << 18) + 37, True, 4, torch.float32),
                                                (1 << 19, True, 8, torch.bfloat16), (4099, False, 3, torch.float32)]):
        for with_nan in (False, True):
            g = torch.Generator().manual_seed(900 + case)
            full = (torch.randn(n, generator=g) * 0.3).to(dt)
            if with_nan:
                full[5] = float("nan")
            shard = torch.tensor_split(full, world)[rank].cuda()
            y, scale = sharded_quantize(shard, bits=bits, plus=plus)
            Q = P.PowerOfTwoPlusQuantizer if plus else P.PowerOfTwoQuantizer
            want = torch.tensor_split(Q.forward(None, full.cuda(), bits=bits), world)[rank]
            same = torch.equal(y.view(torch.int16 if dt != torch.float32 else torch.int32),
                               want.view(torch.int16 if dt != torch.float32 else torch.int32)) or \
                bool((torch.isnan(y) & torch.isnan(want)).all())
            bad += 0 if same else 1
    t = torch.tensor([bad], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"sharded_quantize_cuda_backend_mismatching_cases": int(t.item()), "cases": 8, "world": world}), flush=True)
    ok = ok and t.item() == 0
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
