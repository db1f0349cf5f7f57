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
