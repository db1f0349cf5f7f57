"""TEST INFRASTRUCTURE ONLY -- the reference's hot path restated with stock torch ops.

Why a second oracle: the reference *is* torch code, so restating it op for op with the same ATen
calls gives (a) on CPU, exactly what the reference computes and costs on host cores -- this is the
``cpu_baseline`` / ``bench.py --impl reference`` arm, since the Python reference itself cannot
travel to the GPU box -- and (b) on CUDA, "the reference run on this GPU" (stock ATen/cuDNN
kernels), the same-box comparator the parity tests and the torch_cuda boundary table use.
The numpy oracle (po2_oracle.py) stays the libm-independent checker.

Parity status: PINNED -- tests/test_oracle_golden.py checks these functions bit-exactly against
the vectors generated from the unmodified reference (tests/golden/make_golden.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def quantize_ref(x: torch.Tensor, bits: int = 4, fsr: int = 1, plus: bool = False) -> torch.Tensor:
    """utils/quantizers.py:22-32 (plus=False) / :42-52 (plus=True), one ATen op per reference op."""
    sgn = torch.sign(x)
    s = torch.max(torch.abs(x))
    v = torch.abs(x / s)
    lg = torch.log2(v / 1.5) + 0.5 if plus else torch.log2(v)
    q = torch.clamp(torch.round(lg), fsr - 2 ** (bits - 1), fsr - 1)
    return 2 ** q * sgn * s


class _STE(torch.autograd.Function):
    """utils/quantizers.py:19-56: forward = quantize, backward = identity."""

    @staticmethod
    def forward(ctx, w, bits, plus):
        return quantize_ref(w, bits, 1, plus)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class Po2Oracle:
    """Stands in for PowerOfTwoQuantizer (plus=False) / PowerOfTwoPlusQuantizer (plus=True)."""

    def __init__(self, plus: bool):
        self.plus = plus

    def apply(self, w, bits):
        return _STE.apply(w, bits, self.plus)

    def forward(self, ctx, w, bits=4, fsr=1):
        return quantize_ref(w, bits, fsr, self.plus)


PO2 = Po2Oracle(False)
PO2_PLUS = Po2Oracle(True)


class QuantizedConv2dOracle(nn.Conv2d):
    """models/quantized_conv.py:5-45 restated: quantize the weight every forward, then F.conv2d."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=1, dilation=1,
                 groups=1, bias=False, quantize_fn=None, bits=4):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.quantize_fn = quantize_fn
        self.bits = bits

    def forward(self, input):
        w = self.weight if self.quantize_fn is None else self.quantize_fn.apply(self.weight, self.bits)
        return self._conv_forward(input, w, self.bias)   # models/quantized_conv.py:36,38


def conv2d_ref(x, w_dequant, stride, padding, groups):
    """The conv oracle: F.conv2d on the dequantized fp32 weight (models/quantized_conv.py:36)."""
    return F.conv2d(x, w_dequant, None, stride, padding, 1, groups)


def quantize_model_ref(model: nn.Module, quantizer: Po2Oracle, bits: int) -> float:
    """utils/quantizers.py:139-153."""
    err, numel = 0.0, 0
    with torch.no_grad():
        for _, m in model.named_modules():
            if isinstance(m, QuantizedConv2dOracle):
                for _, p in m.named_parameters():
                    qp = quantizer.forward(None, p, bits=bits)
                    err += torch.sum((qp - p) ** 2)
                    numel += p.numel()
                    p.copy_(qp)
    return (err / numel).item()
