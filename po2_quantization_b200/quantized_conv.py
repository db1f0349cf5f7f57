"""Host-side mirror of the reference's ``models/quantized_conv.py``.

``QuantizedConv2d`` keeps the reference's constructor (note ``padding=1, bias=False`` defaults),
attributes (``quantize_fn``, ``bits``), ``state_dict`` (``{weight}``) and methods, so the
reference's model files build on it unchanged.
"""
import copy

import torch
import torch.nn as nn

from . import ops


class QuantizedConv2d(nn.Conv2d):
    """reference models/quantized_conv.py:5-45"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=1, dilation=1,
                 groups=1, bias=False, quantize_fn=None, bits=4):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.quantize_fn = quantize_fn
        self.bits = bits

    def __deepcopy__(self, memo):
        # test.py:120 deep-copies models; a copied Parameter restarts its version counter, so the
        # (non-persistent) PTQ tag is re-keyed to the copy instead of silently going stale
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        new.__dict__ = {k: copy.deepcopy(v, memo) for k, v in self.__dict__.items()
                        if k not in ("_po2_pack_cache", "_po2_prefetch", "_po2_prefetch_single", "_po2_prefetch_static")}
        tag = self.__dict__.get("_po2_ptq")
        if tag is not None and tag[0] == self.weight._version:
            new.__dict__["_po2_ptq"] = (new.weight._version,) + tuple(new.__dict__["_po2_ptq"][1:])
        return new

    # ---- which inputs the sm_100a conv kernels take.  Anything else goes to nn.Conv2d's own path
    # (cuDNN) -- never silently: _why_not() names the reason and forward() reports it once per reason
    # (ops.note_library_path; an error under PO2_STRICT=1).  Every QuantizedConv2d the reference's model
    # files construct (bias=False, dilation 1, square stride/padding, fp32) is taken.
    def _why_not(self, input):
        if ops.get_conv_mode() == "cudnn":
            return None                                       # the caller asked for cuDNN: not a fallback
        if not input.is_cuda:
            return "CPU input"
        if input.dtype != torch.float32 or self.weight.dtype != torch.float32:
            return f"dtype {input.dtype}/{self.weight.dtype} (fp32 NCHW only)"
        if input.dim() != 4:
            return f"{input.dim()}-D input (batched NCHW only)"
        if self.bias is not None:
            return "bias=True"
        if self.dilation != (1, 1):
            return f"dilation {self.dilation}"
        if self.padding_mode != "zeros" or not isinstance(self.padding, tuple):
            return f"padding {self.padding!r} / padding_mode {self.padding_mode!r}"
        if self.stride[0] != self.stride[1] or self.padding[0] != self.padding[1]:
            return f"non-square stride {self.stride} / padding {self.padding}"
        return ""

    def _po2_conv_ok(self, input) -> bool:
        return self._why_not(input) == ""

    def _library_conv(self, input, weight, why):
        if why:
            ops.note_library_path("qconv:" + why, f"QuantizedConv2d runs nn.Conv2d's own convolution (cuDNN) for this "
                                  f"layer: {why}")
        if input.dim() == 4:
            ops.check_conv_shapes(input.shape, weight.shape, self.groups, self.padding[0] if isinstance(self.padding, tuple) else 0)
        return self._conv_forward(input, weight, self.bias)

    def _po2_conv(self, input, weight, scale):
        return ops.conv2d(input, weight, scale, self.stride[0], self.padding[0], self.groups,
                          ops.COMPUTE[ops.get_conv_mode()])

    def forward(self, input):
        # models/quantized_conv.py:32-38: quantize the weight on every forward (QAT), then conv
        if self.quantize_fn is not None:
            plus = getattr(self.quantize_fn, "_PLUS", None)
            if plus is not None and self._po2_conv_ok(input):
                mode = ops.get_conv_mode()
                if mode in ("tc", "tf32"):
                    # weights already quantized by the multi-tensor prefetch (prefetch.py)?  Then the conv
                    # is a single launch; otherwise remember the input shape for the next prefetch
                    from . import prefetch
                    out = prefetch.try_prefetched_forward(self, input, mode)
                    if out is not None:
                        return out
                    self.__dict__["_po2_xshape"] = tuple(input.shape)
                # PO2 / PO2+: one op = quantizer kernel (which also emits the packed +-2^q tensor-core
                # operand) + conv kernel; the straight-through gradient reaches self.weight
                out, _qw, _scale = ops.qconv2d(input, self.weight, int(self.bits), 1, bool(plus), self.stride[0],
                                               self.padding[0], self.groups, ops.COMPUTE[ops.get_conv_mode()])
                return out
            # lin / lin+ (the conv of their result is nn.Conv2d's: those weights are not on a PO2 grid), or
            # a PO2 layer whose configuration the kernels do not take
            quantized_weight = self.quantize_fn.apply(self.weight, self.bits)
            return self._library_conv(input, quantized_weight, self._why_not(input) if plus is not None else None)
        tag = getattr(self, "_po2_ptq", None)
        if tag is not None and tag[0] == self.weight._version:
            why = self._why_not(input)
            if why:
                return self._library_conv(input, self.weight, why)
            # post-training-quantized weights (quantize_model): already on the grid +-scale*2^q
            mode = ops.get_conv_mode()
            if mode in ("tc", "tf32") and not (torch.is_grad_enabled() and (input.requires_grad or self.weight.requires_grad)):
                # static weights, no autograd: pack the tensor-core operand once per (weight version,
                # input shape) and run each forward as a single launch
                key = (tag[0], tuple(input.shape), input.device, mode)
                cache = self.__dict__.get("_po2_pack_cache")
                if cache is None or cache[0] != key:
                    packed = ops.conv2d_pack(self.weight.detach(), tag[1], tuple(input.shape), self.stride[0],
                                             self.padding[0], self.groups, ops.COMPUTE[mode])
                    cache = (key, packed)
                    self.__dict__["_po2_pack_cache"] = cache
                if cache[1] is not None:
                    K, _, R, S = self.weight.shape
                    return ops.conv2d_packed(input, cache[1], tag[1], K, R, S, self.stride[0], self.padding[0], self.groups,
                                             ops.COMPUTE[mode])
            if why is None:
                return self._conv_forward(input, self.weight, self.bias)
            return self._po2_conv(input, self.weight, tag[1])
        # no quantizer and no PTQ tag: a full-precision layer, i.e. plain nn.Conv2d (models/quantized_conv.py:38)
        return self._conv_forward(input, self.weight, self.bias)

    def forward_with_stats(self, input, stats):
        """forward(input) that, where the layer runs from a prefetched operand on the TMA-fed kernel, also
        accumulates the batch statistics of the norm behind it (stats["sums"], fp64 [sum | sum of squares | ticket])
        in the conv's epilogue; stats["ok"] reports whether it did."""
        stats["ok"] = False
        if self.quantize_fn is not None and getattr(self.quantize_fn, "_PLUS", None) is not None and self._po2_conv_ok(input):
            mode = ops.get_conv_mode()
            if mode == "tf32":
                from . import prefetch
                out = prefetch.try_prefetched_forward(self, input, mode, stats)
                if out is not None:
                    return out
        return self.forward(input)

    def forward_folded(self, input, ep_a, ep_b, residual=None, act=0):
        """Inference forward with a per-out-channel affine (an eval-mode BatchNorm folded in), the residual add
        and the activation in the conv kernel's epilogue: act(conv(input) * ep_a + ep_b + residual).  Returns
        None when this layer / call cannot take the folded path (the caller then runs conv and norm separately):
        autograd active, a configuration the kernels do not take, a QAT layer without a prefetched operand."""
        if torch.is_grad_enabled() and (input.requires_grad or self.weight.requires_grad):
            return None
        mode = ops.get_conv_mode()
        if mode == "cudnn" or self._why_not(input) != "":
            return None
        compute = ops.COMPUTE[mode]
        K, _, R, S = self.weight.shape
        if self.quantize_fn is not None:
            plus = getattr(self.quantize_fn, "_PLUS", None)
            if plus is None or mode not in ("tc", "tf32"):
                return None
            from . import prefetch
            slot = self.__dict__.get("_po2_prefetch")
            if not slot or slot.key != prefetch._layer_key(self, input.shape, mode):
                self.__dict__["_po2_xshape"] = tuple(input.shape)
                return None
            prefetch.wait_for_quantizer(input.device)
            return ops.conv2d_packed_ep(input, slot.packed, slot.scale, K, R, S, self.stride[0], self.padding[0], self.groups,
                                        compute, ep_a, ep_b, residual, int(act))
        tag = getattr(self, "_po2_ptq", None)
        if tag is None or tag[0] != self.weight._version:
            return None                                          # a full-precision layer: plain nn.Conv2d
        if mode in ("tc", "tf32"):
            key = (tag[0], tuple(input.shape), input.device, mode)
            cache = self.__dict__.get("_po2_pack_cache")
            if cache is None or cache[0] != key:
                packed = ops.conv2d_pack(self.weight.detach(), tag[1], tuple(input.shape), self.stride[0],
                                         self.padding[0], self.groups, compute)
                cache = (key, packed)
                self.__dict__["_po2_pack_cache"] = cache
            if cache[1] is not None:
                return ops.conv2d_packed_ep(input, cache[1], tag[1], K, R, S, self.stride[0], self.padding[0], self.groups,
                                            compute, ep_a, ep_b, residual, int(act))
        return ops.conv2d_ep(input, self.weight.detach(), tag[1], self.stride[0], self.padding[0], self.groups, compute,
                             ep_a, ep_b, residual, int(act))

    def get_quantization_error(self):
        """models/quantized_conv.py:40-45: (sum((Q(w) - w)^2), numel); (0, numel) without a quantizer.

        For PO2 / PO2+ on CUDA the sum is the fused fp64 output of the quantizer kernel itself: either the
        value the multi-tensor prefetch already produced for the current weight version (no launch at
        all -- what the models' per-epoch error walkers, train.py:106, then read for all layers), or one
        launch of the single-tensor kernel.  The reference sums in fp32 with 12 ATen launches; the
        value returned here is the fp64 sum rounded once to fp32 (a detached 0-dim tensor)."""
        if self.quantize_fn is not None:
            plus = getattr(self.quantize_fn, "_PLUS", None)
            w = self.weight
            if plus is not None and w.is_cuda and w.dtype in ops._DT and w.numel() > 0:
                slot = self.__dict__.get("_po2_prefetch")
                if (slot and slot.sse is not None and slot.key is not None and slot.key[0] == w._version
                        and slot.key[3] == int(self.bits) and slot.key[4] == bool(plus)):
                    from . import prefetch
                    prefetch.wait_for_quantizer(w.device)
                    return slot.sse.to(torch.float32), w.numel()
                sse = ops.quantize_full(w.detach(), int(self.bits), 1, bool(plus))[4]
                return sse.to(torch.float32), w.numel()
            quantized_weight = self.quantize_fn.apply(self.weight, self.bits)
            return torch.sum((quantized_weight - self.weight) ** 2), self.weight.numel()
        return 0, self.weight.numel()
