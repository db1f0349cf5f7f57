"""world_size-2 gloo tests (CPU) of the N>1 host logic: the sharded quantizer's single exchange step
and DDP-wrappability of the module.  The kernels are replaced by the oracle here (test fake backend);
the GPU suite covers the kernels themselves."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_backend():
    from oracle import po2_oracle as O

    def absmax(x):
        return torch.tensor(float(np.max(np.abs(x.numpy()))) if not torch.isnan(x).any() else float("nan"),
                            dtype=torch.float32)

    def quantize(x, scale, bits, fsr, plus):
        # quantize against an externally supplied scale: plant it so the oracle derives the same one
        xs = np.concatenate([x.numpy().ravel(), np.array([scale.item()], np.float32)])
        return torch.from_numpy(O.quantize(xs, bits, fsr, plus)[:-1].reshape(x.shape))
    return absmax, quantize


def _worker(rank, world, port, with_nan, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from po2_quantization_b200.distributed import sharded_quantize
        g = torch.Generator().manual_seed(77)
        full = torch.randn(4096 + 37, generator=g) * 0.3
        if with_nan:
            full[5] = float("nan")
        shards = torch.tensor_split(full, world)
        y, scale = sharded_quantize(shards[rank], bits=4, plus=True, _backend=_oracle_backend())
        torch.save({"y": y, "scale": scale}, os.path.join(out_dir, f"r{rank}.pt"))
        # DDP wraps the module (fp32 parameter, state_dict == {weight}) without touching the kernels
        from po2_quantization_b200 import QuantizedConv2d
        m = QuantizedConv2d(4, 4, 3, quantize_fn=None)
        ddp = torch.nn.parallel.DistributedDataParallel(m)
        out = ddp(torch.randn(2, 4, 8, 8))           # quantize_fn None on CPU -> nn.Conv2d's own path
        out.sum().backward()
        grads = [torch.zeros_like(m.weight.grad) for _ in range(world)]
        dist.all_gather(grads, m.weight.grad)
        assert all(torch.equal(grads[0], gk) for gk in grads), "DDP did not average the gradients"
        assert [k for k in ddp.state_dict()] == ["module.weight"]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("with_nan", [False, True])
def test_sharded_quantize_matches_whole_tensor(tmp_path, with_nan):
    from oracle import po2_oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), with_nan, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(77)
    full = torch.randn(4096 + 37, generator=g) * 0.3
    if with_nan:
        full[5] = float("nan")
    ref = O.quantize(full.numpy(), 4, 1, True)
    got = torch.cat([torch.load(os.path.join(tmp_path, f"r{r}.pt"))["y"] for r in range(world)]).numpy()
    scales = [torch.load(os.path.join(tmp_path, f"r{r}.pt"))["scale"].item() for r in range(world)]
    if with_nan:
        assert all(np.isnan(s) for s in scales) and np.isnan(got).all() and np.isnan(ref).all()
    else:
        assert scales[0] == scales[1] == float(np.max(np.abs(full.numpy())))
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def _sharded_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from po2_quantization_b200.distributed import BatchSharded
        torch.manual_seed(100 + rank)                       # replicas start DIFFERENT: the wrapper must broadcast rank 0's
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.BatchNorm2d(4), torch.nn.Flatten(),
                                  torch.nn.Linear(4 * 6 * 6, 5))
        model = BatchSharded(net)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(8, 3, 6, 6, generator=g)
        t = torch.randint(0, 5, (8,), generator=g)
        sl = slice(rank * 4, rank * 4 + 4)
        loss = torch.nn.functional.cross_entropy(model(x[sl]), t[sl])
        loss.backward()
        n = model.average_gradients()
        assert n == len(list(net.parameters()))
        torch.save({"state": {k: v.clone() for k, v in net.state_dict().items()},
                    "grads": [p.grad.clone() for p in net.parameters()]}, os.path.join(out_dir, f"s{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_batch_sharded_broadcasts_and_averages_like_ddp(tmp_path):
    """BatchSharded == DistributedDataParallel arithmetic: rank 0's parameters everywhere, gradients
    averaged over ranks (checked against a single-process run over the two shards)."""
    world = 2
    mp.spawn(_sharded_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"s{k}.pt")) for k in range(world)]
    for a, b in zip(r[0]["grads"], r[1]["grads"]):
        assert torch.equal(a, b)
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.BatchNorm2d(4), torch.nn.Flatten(),
                              torch.nn.Linear(4 * 6 * 6, 5))
    for k, v in net.state_dict().items():
        if "running" not in k and "num_batches" not in k:
            assert torch.equal(v, r[1]["state"][k]), k          # rank 1 holds rank 0's initial parameters
    g = torch.Generator().manual_seed(5)
    x = torch.randn(8, 3, 6, 6, generator=g)
    t = torch.randint(0, 5, (8,), generator=g)
    ref = [torch.zeros_like(p) for p in net.parameters()]
    for k in range(world):
        net.zero_grad()
        net.load_state_dict({kk: vv for kk, vv in r[0]["state"].items()})
        torch.nn.functional.cross_entropy(net(x[k * 4:k * 4 + 4]), t[k * 4:k * 4 + 4]).backward()
        for acc, p in zip(ref, net.parameters()):
            acc += p.grad / world
    for a, b in zip(r[0]["grads"], ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
