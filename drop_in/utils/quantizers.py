"""Drop-in replacement for the reference's utils/quantizers.py."""
from po2_quantization_b200.quantizers import (  # noqa: F401
    LinearPowerOfTwoPlusQuantizer, LinearPowerOfTwoQuantizer, PowerOfTwoPlusQuantizer,
    PowerOfTwoQuantizer, quantize_model, quantizer_dict)
