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
    is a few coalesced NCCL all-reduce(AVG) calls over the gradient tensors where they lie.  DDP's
    reducer instead launches one scale-and-copy kernel per parameter into its buckets (171 launches
    per ResNet-56 step, ~0.4 ms on a B200, as long as the whole gradient all-reduce itself), which is
    what this avoids.  With ``overlap=True`` (CUDA) the parameters are split into ``buckets`` groups
    in backward order; a group's all-reduce is issued from a post-accumulate-grad hook as soon as its
    last gradient exists, on a side stream, so it runs under the rest of the backward pass (also inside
    a CUDA-graph capture, where the side stream becomes a parallel branch).  ``average_gradients()``
    after ``backward()`` reduces whatever is left and joins the side stream.

    ``overlap`` defaults to False since round 2: measured on the ResNet-56 step (3.4 MB of gradients), ONE coalesced
    all-reduce behind backward beats the overlapped buckets -- 3.65 vs 3.80 ms per step at 2 GPUs, 3.80 vs 4.04 ms at
    8 -- because the NCCL kernels running under backward take SMs from kernels that are sized one or two CTAs per
    SM (persistent tcgen05 convs, cooperative norm kernels), which costs more than the exposed tail of a 3.4 MB
    all-reduce over NVLink.  Models with far larger gradients may prefer ``overlap=True``."""

    def __init__(self, module: torch.nn.Module, group: Optional[dist.ProcessGroup] = None, overlap: bool = False,
                 buckets: int = 4):
        super().__init__()
        self.module = module
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._buckets, self._pending, self._bucket_of, self._comm = [], [], {}, None
        if self.world > 1:
            src = dist.get_global_rank(group, 0) if group is not None else 0
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=src, group=group)
            params = [p for p in module.parameters() if p.requires_grad]
            if overlap and params and params[0].is_cuda and dist.get_backend(group) == "nccl":
                self._comm = torch.cuda.Stream(params[0].device)
                total = sum(p.numel() for p in params)
                acc, cur = 0, []
                for p in reversed(params):                      # gradients arrive in roughly this order
                    cur.append(p)
                    acc += p.numel()
                    if acc * buckets >= total * (len(self._buckets) + 1) and len(self._buckets) < buckets - 1:
                        self._buckets.append(cur)
                        cur = []
                if cur:
                    self._buckets.append(cur)
                for bi, bucket in enumerate(self._buckets):
                    for p in bucket:
                        self._bucket_of[id(p)] = bi
                        p.register_post_accumulate_grad_hook(self._on_grad)
                self._pending = [len(b) for b in self._buckets]

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # ---- gradient exchange ---------------------------------------------------------------------------
    def _all_reduce(self, grads) -> None:
        if dist.get_backend(self.group) == "nccl":
            with dist._coalescing_manager(group=self.group, device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
        else:                                   # gloo (CPU tests): no AVG, no coalescing
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            torch._foreach_div_(grads, float(self.world))

    def _launch_bucket(self, bi: int) -> int:
        grads = [p.grad for p in self._buckets[bi] if p.grad is not None]
        self._pending[bi] = -1                  # done for this step
        if grads:
            cur = torch.cuda.current_stream(grads[0].device)
            self._comm.wait_stream(cur)         # the gradients of this bucket are complete on the compute stream
            with torch.cuda.stream(self._comm):
                self._all_reduce(grads)
        return len(grads)

    def _on_grad(self, p) -> None:
        bi = self._bucket_of[id(p)]
        if self._pending[bi] < 0:
            # this bucket was already all-reduced for the current step: a second backward() (gradient
            # accumulation) would add un-averaged local gradients on top of averaged ones and race with the
            # all-reduce still running on the side stream -- ranks would silently diverge
            raise RuntimeError("BatchSharded: backward() ran again before average_gradients(); call "
                               "average_gradients() after every backward (gradient accumulation is not supported "
                               "with overlap=True -- construct BatchSharded(..., overlap=False) for that)")
        if self._pending[bi] > 0:
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch_bucket(bi)

    def average_gradients(self) -> int:
        """Average every existing .grad over the group (what was not already reduced from the hooks);
        returns how many gradient tensors this step exchanged."""
        if self.world <= 1:
            return sum(1 for p in self.module.parameters() if p.grad is not None)
        from . import batchnorm
        for ex in batchnorm._exchanges.values():             # a timed-out SyncBatchNorm exchange raises here
            if ex is not None:
                ex.check()
        if self._comm is None:
            grads = [p.grad for p in self.module.parameters() if p.grad is not None]
            if grads:
                self._all_reduce(grads)
            return len(grads)
        n = 0
        for bi, bucket in enumerate(self._buckets):
            if self._pending[bi] >= 0:          # a parameter of this bucket got no gradient this step
                n += self._launch_bucket(bi)
            else:
                n += sum(1 for p in bucket if p.grad is not None)
        dev = self._buckets[0][0].device
        torch.cuda.current_stream(dev).wait_stream(self._comm)
        self._pending = [len(b) for b in self._buckets]
        return n
