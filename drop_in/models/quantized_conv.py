"""Drop-in replacement for the reference's models/quantized_conv.py."""
from po2_quantization_b200.quantized_conv import QuantizedConv2d  # noqa: F401
