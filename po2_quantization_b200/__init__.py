"""po2_quantization_b200 -- B200 (sm_100a) implementation of the po2_quantization hot path.

Public surface mirrors the reference (utils/quantizers.py, models/quantized_conv.py):
``PowerOfTwoQuantizer``, ``PowerOfTwoPlusQuantizer``, ``LinearPowerOfTwoQuantizer``,
``LinearPowerOfTwoPlusQuantizer``, ``quantize_model``, ``quantizer_dict``, ``QuantizedConv2d``.
"""
from . import _lib, ops  # noqa: F401
from .batchnorm import FusedSyncBatchNorm  # noqa: F401
from .prefetch import enable_weight_prefetch, prefetch_weights  # noqa: F401
from .ops import get_log2_flavor, set_log2_flavor  # noqa: F401
from .quantized_conv import QuantizedConv2d  # noqa: F401
from .quantizers import (LinearPowerOfTwoPlusQuantizer, LinearPowerOfTwoQuantizer,  # noqa: F401
                         PowerOfTwoPlusQuantizer, PowerOfTwoQuantizer, model_quantization_error, quantize_model,
                         quantizer_dict)
from .fold import FoldedConvBN, conv_bn_act, fold_conv_bn, fuse_batchnorm, invalidate_caches  # noqa: F401
from . import optim  # noqa: F401
from .stem import StemConv2d, accelerate_stem  # noqa: F401
from .checkpoint import load_packed_checkpoint, pack_state_dict, save_packed_checkpoint, unpack_state_dict  # noqa: F401

__version__ = "0.1.0"
