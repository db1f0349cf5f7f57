"""Synthetic workloads for bench.py and the parity tests: the layer graphs BASELINE.json's configs
name, built on this repo's ``QuantizedConv2d``.  They are harnesses
around the hot path, not part of it: BN / ReLU / Linear are stock torch modules, and parameter
names follow the reference's models so its checkpoints load."""
from .resnet_cifar import resnet_cifar  # noqa: F401
from .mobilenet_cifar import mobilenet_v2_cifar  # noqa: F401
from .mobilevit import mobilevit_xs  # noqa: F401
