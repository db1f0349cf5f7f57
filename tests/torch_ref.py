"""TEST INFRASTRUCTURE: the reference's op sequence restated with stock torch ops.

Runs on whatever device the input lives on, so on the B200 box it is "the reference on a GPU"
(stock ATen CUDA kernels + CUDA libm) and on CPU it is "the reference on a CPU".
Follows utils/quantizers.py:22-32 / :42-52 and models/quantized_conv.py:32-45.
"""
import torch


def quantize_ref(x: torch.Tensor, bits: int = 4, fsr: int = 1, plus: bool = False) -> torch.Tensor:
    sgn = torch.sign(x)
    s = torch.max(torch.abs(x))
    v = torch.abs(x / s)
    lg = torch.log2(v / 1.5) + 0.5 if plus else torch.log2(v)
    q = torch.clamp(torch.round(lg), fsr - 2 ** (bits - 1), fsr - 1)
    return 2 ** q * sgn * s


def quantize_ref_parts(x, bits=4, fsr=1, plus=False):
    sgn = torch.sign(x)
    s = torch.max(torch.abs(x))
    v = torch.abs(x / s)
    lg = torch.log2(v / 1.5) + 0.5 if plus else torch.log2(v)
    q = torch.clamp(torch.round(lg), fsr - 2 ** (bits - 1), fsr - 1)
    return 2 ** q * sgn * s, q, sgn, s
