"""The full-precision stem conv of the reference's models (models/resnet.py:99-102 ``nn.Conv2d(3, 16, 3, 1, 1,
bias=False)``, "first conv unquantized"): its forward stays ``F.conv2d`` (cuDNN, the reference's own arithmetic), its
backward -- only a weight gradient, the images need none -- runs on ``po2_conv2d_stem_wgrad`` instead of cuDNN's
40 us wgrad kernel.  ``StemConv2d`` is an ``nn.Conv2d`` subclass (same constructor, ``state_dict``, initialisation
loops that test ``isinstance(m, nn.Conv2d)``); ``accelerate_stem(model)`` re-classes the qualifying plain ``nn.Conv2d``
instances of an unmodified model in place, the way ``fuse_batchnorm`` does for the norms."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops


def _takes(x, w, stride, pad, dilation, groups) -> bool:
    if not (x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and x.dim() == 4):
        return False
    if tuple(stride) != (1, 1) or tuple(pad) != (1, 1) or tuple(dilation) != (1, 1) or groups != 1:
        return False
    B, C, H, W_ = x.shape
    K, _, R, S = w.shape
    return int(_lib.load().po2_conv2d_stem_wgrad_workspace(B, C, H, W_, K, R, S, 1, 1, 1)) > 0


class _StemConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return F.conv2d(x, w, None, 1, 1)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = gw = None
        if ctx.needs_input_grad[0]:                              # not the stem's case: the library path, both gradients
            gx, gw, _ = torch.ops.aten.convolution_backward(g, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                            [True, ctx.needs_input_grad[1], False])
            return gx, gw
        if ctx.needs_input_grad[1]:
            g = g.contiguous()
            xc = x.contiguous()
            B, C, H, W_ = xc.shape
            K = w.shape[0]
            lib = _lib.load()
            with torch.cuda.device(x.device):
                gw = torch.empty_like(w, memory_format=torch.contiguous_format)
                ws = torch.empty(int(lib.po2_conv2d_stem_wgrad_workspace(B, C, H, W_, K, 3, 3, 1, 1, 1)), dtype=torch.uint8,
                                 device=x.device)
                ops.LAUNCHES += 2
                _lib.check(lib.po2_conv2d_stem_wgrad(g.data_ptr(), xc.data_ptr(), gw.data_ptr(), B, C, H, W_, K, 3, 3, 1, 1, 1,
                                                     ws.data_ptr(), ws.numel(), ops._stream_ptr(x.device)),
                           "po2_conv2d_stem_wgrad")
        return gx, gw


class StemConv2d(nn.Conv2d):
    """nn.Conv2d whose weight gradient runs on the small-C kernel when the layer is a 3x3 / stride 1 / pad 1 conv with
    at most 4 input channels on fp32 CUDA tensors; everything else is nn.Conv2d.forward."""

    def forward(self, input):
        if (self.bias is None and self.padding_mode == "zeros" and torch.is_grad_enabled() and self.weight.requires_grad
                and _takes(input, self.weight, self.stride, self.padding, self.dilation, self.groups)):
            return _StemConv.apply(input, self.weight)
        return super().forward(input)


def accelerate_stem(model: nn.Module) -> int:
    """Re-class the plain ``nn.Conv2d`` instances that ``StemConv2d`` can take (3x3, stride 1, pad 1, <= 4 input
    channels, no bias) in place; parameters and ``state_dict`` keys do not change.  Returns how many."""
    n = 0
    for m in model.modules():
        if (type(m) is nn.Conv2d and m.in_channels <= 4 and m.kernel_size == (3, 3) and m.stride == (1, 1)
                and m.padding == (1, 1) and m.dilation == (1, 1) and m.groups == 1 and m.bias is None):
            m.__class__ = StemConv2d
            n += 1
    return n
