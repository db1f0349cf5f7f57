"""torch.library custom ops over the C ABI.

Torch's role here is plumbing: it owns device memory and the current stream.  Every op enqueues
hand-written sm_100a kernels from libpo2b200.so on ``torch.cuda.current_stream()`` and is CUDA
graph capturable.  CPU tensors are rejected -- there is no fallback path.
"""
import os
from typing import Optional, Tuple

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}
_FLAVORS = {"ieee": _lib.FLAVOR_IEEE, "torch_cpu": _lib.FLAVOR_IEEE, "torch_cuda": _lib.FLAVOR_TORCH_CUDA}
_flavor = _FLAVORS[os.environ.get("PO2_LOG2_FLAVOR", "ieee")]


def set_log2_flavor(name: str) -> None:
    """Which float-log2 the rounding boundaries reproduce: "ieee" (correctly rounded == torch CPU,
    the default and what the oracle uses) or "torch_cuda" (torch's CUDA kernels)."""
    global _flavor
    if name not in _FLAVORS:
        raise ValueError(f"unknown log2 flavor {name!r}; choose from {sorted(_FLAVORS)}")
    if _FLAVORS[name] == _lib.FLAVOR_TORCH_CUDA and not _lib.load().po2_have_torch_cuda_table():
        raise _lib.Po2Error("the torch_cuda boundary table has not been scanned into this build")
    _flavor = _FLAVORS[name]


def get_log2_flavor() -> str:
    return "torch_cuda" if _flavor == _lib.FLAVOR_TORCH_CUDA else "ieee"


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _lib.Po2Error(f"{what}: po2_quantization_b200 runs on CUDA (sm_100a) only -- got a "
                            f"{t.device.type} tensor and there is no CPU fallback")
    if t.dtype not in _DT:
        raise TypeError(f"{what}: unsupported dtype {t.dtype} (float32, bfloat16, float16)")


_workspaces = {}

# number of kernels of libpo2b200.so launched through this module (bench.py's `gpu_launches`)
LAUNCHES = 0


def conv_backend_name() -> str:
    """Which kernel QuantizedConv2d's convolution runs on (reported by bench.py)."""
    return "cudnn (torch F.conv2d on the po2-quantized weight)"


def _workspace(device: torch.device) -> torch.Tensor:
    """Zero-initialised scratch, one per (device, stream); the kernels leave it zeroed."""
    stream = torch.cuda.current_stream(device)
    key = (device.index, stream.cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(int(_lib.load().po2_workspace_bytes()), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _code_bytes(n: int, bits: int) -> int:
    return (n + 1) // 2 if bits <= 4 else n


# ------------------------------------------------------------------------------------------------
# raw launchers (no autograd, no dispatcher) -- also what bench.py times
# ------------------------------------------------------------------------------------------------
def absmax_out(x: torch.Tensor, scale: torch.Tensor) -> None:
    global LAUNCHES
    lib = _lib.load()
    LAUNCHES += 1
    _lib.check(lib.po2_absmax(x.data_ptr(), x.numel(), _DT[x.dtype], scale.data_ptr(),
                              _workspace(x.device).data_ptr(), _stream_ptr(x.device)), "po2_absmax")


def quantize_out(x, y, scale, bits, fsr, plus, codes=None, zero_count=None, sse=None, flavor=None):
    global LAUNCHES
    lib = _lib.load()
    LAUNCHES += 1
    _lib.check(lib.po2_quantize(
        x.data_ptr(), y.data_ptr(), codes.data_ptr() if codes is not None else None,
        zero_count.data_ptr() if zero_count is not None else None,
        sse.data_ptr() if sse is not None else None, scale.data_ptr(), x.numel(), _DT[x.dtype],
        bits, fsr, int(plus), _flavor if flavor is None else flavor, _stream_ptr(x.device)), "po2_quantize")


def quantize_fused_out(x, y, scale, bits, fsr, plus, codes=None, zero_count=None, sse=None, flavor=None):
    global LAUNCHES
    lib = _lib.load()
    aligned = (x.data_ptr() | y.data_ptr()) % 16 == 0
    LAUNCHES += lib.po2_quantize_fused_launches(x.numel(), _DT[x.dtype]) if aligned else 2
    _lib.check(lib.po2_quantize_fused(
        x.data_ptr(), y.data_ptr(), codes.data_ptr() if codes is not None else None,
        zero_count.data_ptr() if zero_count is not None else None,
        sse.data_ptr() if sse is not None else None, scale.data_ptr(), x.numel(), _DT[x.dtype],
        bits, fsr, int(plus), _flavor if flavor is None else flavor,
        _workspace(x.device).data_ptr(), _stream_ptr(x.device)), "po2_quantize_fused")


# ------------------------------------------------------------------------------------------------
# custom ops
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("po2::quantize", mutates_args=(), device_types="cuda")
def quantize(x: torch.Tensor, bits: int, fsr: int, plus: bool) -> torch.Tensor:
    """y = 2^clamp(round(log2|x/s|)) * sign(x) * s, s = max|x| -- utils/quantizers.py:21-32, 41-52."""
    _require_cuda(x, "po2::quantize")
    x = x.contiguous()
    y = torch.empty_like(x)
    if x.numel() == 0:
        raise RuntimeError("max(): Expected reduction dim to be specified for input.numel() == 0")
    with torch.cuda.device(x.device):
        scale = torch.empty((), dtype=torch.float32, device=x.device)
        quantize_fused_out(x, y, scale, bits, fsr, plus)
    return y


@quantize.register_fake
def _(x, bits, fsr, plus):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


def _quantize_bwd(ctx, grad):
    # straight-through estimator: utils/quantizers.py:34-36 returns grad_output itself
    return grad, None, None, None


quantize.register_autograd(_quantize_bwd)


@torch.library.custom_op("po2::quantize_full", mutates_args=(), device_types="cuda")
def quantize_full(x: torch.Tensor, bits: int, fsr: int, plus: bool
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(y, packed codes, scale, zero_count, sse): everything one pass over x can produce."""
    _require_cuda(x, "po2::quantize_full")
    x = x.contiguous()
    if x.numel() == 0:
        raise RuntimeError("max(): Expected reduction dim to be specified for input.numel() == 0")
    with torch.cuda.device(x.device):
        y = torch.empty_like(x)
        codes = torch.empty(_code_bytes(x.numel(), bits), dtype=torch.uint8, device=x.device)
        scale = torch.empty((), dtype=torch.float32, device=x.device)
        zero_count = torch.zeros((), dtype=torch.int32, device=x.device)
        sse = torch.zeros((), dtype=torch.float64, device=x.device)
        quantize_fused_out(x, y, scale, bits, fsr, plus, codes, zero_count, sse)
    return y, codes, scale, zero_count, sse


@quantize_full.register_fake
def _(x, bits, fsr, plus):
    n = x.numel()
    return (torch.empty_like(x, memory_format=torch.contiguous_format),
            x.new_empty(_code_bytes(n, bits), dtype=torch.uint8), x.new_empty((), dtype=torch.float32),
            x.new_empty((), dtype=torch.int32), x.new_empty((), dtype=torch.float64))


@torch.library.custom_op("po2::dequantize", mutates_args=(), device_types="cuda")
def dequantize(codes: torch.Tensor, scale: torch.Tensor, numel: int, bits: int, fsr: int,
               dtype: torch.dtype) -> torch.Tensor:
    """+-2^q * scale from packed sign+exponent codes."""
    if not codes.is_cuda:
        raise _lib.Po2Error("po2::dequantize: CUDA only")
    if dtype not in _DT:
        raise TypeError(f"po2::dequantize: unsupported dtype {dtype}")
    y = torch.empty(numel, dtype=dtype, device=codes.device)
    with torch.cuda.device(codes.device):
        _lib.check(_lib.load().po2_dequantize(codes.data_ptr(), scale.data_ptr(), y.data_ptr(), numel,
                                              _DT[dtype], bits, fsr, _stream_ptr(codes.device)),
                   "po2_dequantize")
    return y


@dequantize.register_fake
def _(codes, scale, numel, bits, fsr, dtype):
    return codes.new_empty(numel, dtype=dtype)


@torch.library.custom_op("po2::ste_backward", mutates_args=("grad_input",), device_types="cuda")
def ste_backward(grad_output: torch.Tensor, grad_input: torch.Tensor, accumulate: bool) -> None:
    """grad_input = grad_output (or += when accumulate) -- utils/quantizers.py:34-36, 54-56."""
    _require_cuda(grad_output, "po2::ste_backward")
    if grad_input.dtype != grad_output.dtype or grad_input.numel() != grad_output.numel():
        raise ValueError("po2::ste_backward: grad_input must match grad_output")
    if not (grad_output.is_contiguous() and grad_input.is_contiguous()):
        raise ValueError("po2::ste_backward: contiguous tensors only")
    with torch.cuda.device(grad_output.device):
        _lib.check(_lib.load().po2_ste_backward(grad_output.data_ptr(), grad_input.data_ptr(),
                                                grad_output.numel(), _DT[grad_output.dtype],
                                                int(accumulate), _stream_ptr(grad_output.device)),
                   "po2_ste_backward")
