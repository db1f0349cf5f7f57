"""Host-side mirror of the reference's ``models/quantized_conv.py``.

``QuantizedConv2d`` keeps the reference's constructor (note ``padding=1, bias=False`` defaults),
attributes (``quantize_fn``, ``bits``), ``state_dict`` (``{weight}``) and methods, so the
reference's model files build on it unchanged.
"""
import torch
import torch.nn as nn


class QuantizedConv2d(nn.Conv2d):
    """reference models/quantized_conv.py:5-45"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=1, dilation=1,
                 groups=1, bias=False, quantize_fn=None, bits=4):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.quantize_fn = quantize_fn
        self.bits = bits

    def forward(self, input):
        # models/quantized_conv.py:32-38: quantize the weight on every forward (QAT), then conv
        if self.quantize_fn is not None:
            quantized_weight = self.quantize_fn.apply(self.weight, self.bits)
            return self._conv_forward(input, quantized_weight, self.bias)
        return self._conv_forward(input, self.weight, self.bias)

    def get_quantization_error(self):
        # models/quantized_conv.py:40-45
        if self.quantize_fn is not None:
            quantized_weight = self.quantize_fn.apply(self.weight, self.bits)
            return torch.sum((quantized_weight - self.weight) ** 2), self.weight.numel()
        return 0, self.weight.numel()
