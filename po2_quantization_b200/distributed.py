"""Multi-GPU pieces of the hot path.

* Quantized-conv inference and QAT shard by batch: one process per GPU, replicated weights, no
  data-path collective of ours (QAT's gradient all-reduce is DistributedDataParallel's NCCL call).
* One huge tensor quantized across ranks has a single exchange step: the per-tensor scale is a
  global max, so each rank reduces its shard locally, the scales meet in one all_reduce(MAX) of a
  single float, and every rank then quantizes its shard against the same scale.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def _cuda_backend() -> Tuple[Callable, Callable]:
    def absmax(x):
        s = torch.empty((), dtype=torch.float32, device=x.device)
        ops.absmax_out(x, s)
        return s

    def quantize(x, scale, bits, fsr, plus):
        y = torch.empty_like(x)
        ops.quantize_out(x, y, scale, bits, fsr, plus)
        return y
    return absmax, quantize


def sharded_quantize(x_shard: torch.Tensor, bits: int = 4, plus: bool = False, fsr: int = 1,
                     group: Optional[dist.ProcessGroup] = None,
                     _backend: Optional[Tuple[Callable, Callable]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Quantize this rank's shard of a tensor that is partitioned across the process group.

    Returns (quantized shard, global scale).  Equivalent to running the quantizer on the
    concatenation of all shards (utils/quantizers.py:21-32 / 41-52) and slicing the result.
    `_backend` (absmax_fn, quantize_fn) exists so the CPU test-suite can drive the exchange logic
    with the oracle; the product path is the sm_100a kernels."""
    if _backend is None:
        ops._require_cuda(x_shard, "sharded_quantize")
        _backend = _cuda_backend()
    absmax, quantize = _backend
    x_shard = x_shard.contiguous()
    scale = absmax(x_shard)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        # NaN must win the reduction like it wins torch.max: MAX over the *bit patterns* of
        # non-negative floats is order-preserving and ranks NaN above +inf
        bits_view = scale.view(torch.int32)
        dist.all_reduce(bits_view, op=dist.ReduceOp.MAX, group=group)
    return quantize(x_shard, scale, bits, fsr, plus), scale


class BatchSharded(torch.nn.Module):
    """Data-parallel QAT wrapper: every rank holds a replica and a
    shard of the batch; after backward the gradients are averaged over the group.

    Same arithmetic as the reference's DistributedDataParallel (train.py:153-155) -- parameters and
    buffers broadcast from rank 0 at construction, gradients averaged every step -- but the exchange
    is ONE coalesced NCCL all-reduce(AVG) over the gradient tensors where they lie, issued by
    ``average_gradients()`` after ``backward()``.  DDP's reducer instead launches one scale-and-copy
    kernel per parameter into its buckets (171 launches per ResNet-56 step, ~0.4 ms on a B200, as
    long as the whole gradient all-reduce itself), which is what this avoids; with 3.4 MB of
    gradients there is nothing worth overlapping with backward."""

    def __init__(self, module: torch.nn.Module, group: Optional[dist.ProcessGroup] = None):
        super().__init__()
        self.module = module
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if self.world > 1:
            src = dist.get_global_rank(group, 0) if group is not None else 0
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=src, group=group)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def average_gradients(self) -> int:
        """all-reduce(AVG) every existing .grad in one coalesced call; returns how many tensors"""
        grads = [p.grad for p in self.module.parameters() if p.grad is not None]
        if self.world <= 1 or not grads:
            return len(grads)
        if dist.get_backend(self.group) == "nccl":
            with dist._coalescing_manager(group=self.group, device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
        else:                                   # gloo (CPU tests): no AVG, no coalescing
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            torch._foreach_div_(grads, float(self.world))
        return len(grads)
