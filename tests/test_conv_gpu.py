"""Quantized-conv forward parity (models/quantized_conv.py:32-38): the sm_100a kernels against the
oracle conv (F.conv2d on the dequantized fp32 weight, evaluated in fp64).
Tolerances (BASELINE.json north_star (b)): rel 1e-2 for the bf16 tensor-core path, 1e-5 for the
fp32-accumulate path; rel = max|out - ref| / max|ref|."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import po2_quantization_b200  # noqa: E402,F401  (registers torch.ops.po2.*)

TOL_TC = 1e-2
TOL_FP32 = 1e-5

# (name, B, C, H, W, K, k, stride, pad, groups) -- SURVEY.md section 8a layer tables
RESNET = [
    ("r_16_16_3s1", 128, 16, 32, 32, 16, 3, 1, 1, 1),
    ("r_16_32_3s2", 128, 16, 32, 32, 32, 3, 2, 1, 1),
    ("r_16_32_1s2", 128, 16, 32, 32, 32, 1, 2, 0, 1),
    ("r_32_32_3s1", 128, 32, 16, 16, 32, 3, 1, 1, 1),
    ("r_32_64_3s2", 128, 32, 16, 16, 64, 3, 2, 1, 1),
    ("r_32_64_1s2", 128, 32, 16, 16, 64, 1, 2, 0, 1),
    ("r_64_64_3s1", 128, 64, 8, 8, 64, 3, 1, 1, 1),
]
MOBILENET = [
    ("m_pw_32_16", 128, 32, 16, 16, 16, 1, 1, 0, 1),
    ("m_pw_16_96", 128, 16, 16, 16, 96, 1, 1, 0, 1),
    ("m_pw_96_24", 128, 96, 8, 8, 24, 1, 1, 0, 1),
    ("m_pw_24_144", 128, 24, 8, 8, 144, 1, 1, 0, 1),
    ("m_pw_144_32", 128, 144, 4, 4, 32, 1, 1, 0, 1),
    ("m_pw_64_384", 128, 64, 2, 2, 384, 1, 1, 0, 1),
    ("m_pw_576_160", 128, 576, 1, 1, 160, 1, 1, 0, 1),
    ("m_pw_160_960", 128, 160, 1, 1, 960, 1, 1, 0, 1),
    ("m_pw_960_320", 128, 960, 1, 1, 320, 1, 1, 0, 1),
    ("m_dw_32_s1", 128, 32, 16, 16, 32, 3, 1, 1, 32),
    ("m_dw_96_s2", 128, 96, 16, 16, 96, 3, 2, 1, 96),
    ("m_dw_576_s2", 128, 576, 2, 2, 576, 3, 2, 1, 576),
    ("m_dw_960_s1", 128, 960, 1, 1, 960, 3, 1, 1, 960),
]
ODD = [
    ("odd_small_batch", 3, 16, 32, 32, 16, 3, 1, 1, 1),
    ("odd_rect", 5, 48, 14, 9, 40, 3, 1, 1, 1),
    ("odd_1x1_big", 2, 128, 56, 56, 64, 1, 1, 0, 1),
    ("odd_grouped", 4, 32, 8, 8, 64, 3, 1, 1, 4),
    ("odd_5x5", 2, 8, 12, 12, 8, 5, 1, 2, 1),
    ("odd_mvit_128_64", 4, 128, 28, 28, 64, 3, 1, 1, 1),
]
ALL = RESNET + MOBILENET + ODD


def _make(case, bits=4, plus=True, seed=0):
    name, B, C, H, W, K, k, stride, pad, groups = case
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    w = torch.randn(K, C // groups, k, k, device="cuda", generator=g) * 0.1
    y, codes, scale, zc, sse = torch.ops.po2.quantize_full(w, bits, 1, plus)
    return x, y, codes, scale


def _ref(x, y, stride, pad, groups):
    return F.conv2d(x.double(), y.double(), None, stride, pad, 1, groups)


def _rel(out, ref):
    return ((out.double() - ref).abs().max() / ref.abs().max()).item()


TOL_TF32 = 2e-3


@pytest.mark.parametrize("compute", [0, 2], ids=["bf16", "tf32"])
@pytest.mark.parametrize("case", ALL, ids=[c[0] for c in ALL])
def test_conv_tensor_core_path(case, compute):
    name, B, C, H, W, K, k, stride, pad, groups = case
    x, y, codes, scale = _make(case)
    out = torch.ops.po2.conv2d(x, y, scale, stride, pad, groups, compute)
    ref = _ref(x, y, stride, pad, groups)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert _rel(out, ref) < (TOL_TC if compute == 0 else TOL_TF32), (name, _rel(out, ref))
    # relative RMS error is the tighter statement: rounded activations, exact weights, fp32 accumulate
    rms = ((out.double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert rms < (4e-3 if compute == 0 else 5e-4), (name, rms)


@pytest.mark.parametrize("case", ALL, ids=[c[0] for c in ALL])
def test_conv_fp32_accumulate_path(case):
    name, B, C, H, W, K, k, stride, pad, groups = case
    x, y, codes, scale = _make(case)
    out = torch.ops.po2.conv2d(x, y, scale, stride, pad, groups, 1)
    assert _rel(out, _ref(x, y, stride, pad, groups)) < TOL_FP32, name


@pytest.mark.parametrize("case", [RESNET[0], RESNET[6], MOBILENET[4], MOBILENET[9]], ids=lambda c: c[0])
@pytest.mark.parametrize("bits", [4, 8])
def test_conv_from_packed_codes_equals_fp32_weights(case, bits):
    """The packed sign+exponent codes carry exactly the information the conv needs."""
    from po2_quantization_b200 import _lib, ops
    name, B, C, H, W, K, k, stride, pad, groups = case
    x, y, codes, scale = _make(case, bits=bits)
    a = torch.ops.po2.conv2d(x, y, scale, stride, pad, groups, 0)
    b = torch.empty_like(a)
    ops.conv2d_out(x, codes, scale, b, stride, pad, groups, 0, w_format=_lib.W_CODES, bits=bits, fsr=1,
                   wshape=tuple(y.shape))
    if groups == 1:
        assert torch.equal(a, b), name          # same exact bf16 operand either way
    else:
        assert _rel(b, _ref(x, y, stride, pad, groups)) < TOL_FP32


@pytest.mark.parametrize("case", [RESNET[0], RESNET[3], RESNET[6], MOBILENET[1], MOBILENET[4], ODD[1]], ids=lambda c: c[0])
def test_conv_dgrad_tensor_core_path(case):
    """Data gradient on the tcgen05 kernel (transposed, rotated PO2 weights; grad rounded to bf16)
    against fp64 conv_transpose: rel 1e-2 like the forward."""
    from po2_quantization_b200 import ops
    name, B, C, H, W, K, k, stride, pad, groups = case
    x, y, codes, scale = _make(case)
    gout = torch.randn(B, K, H, W, device="cuda")
    gx = torch.empty_like(x)
    assert ops.conv2d_dgrad_out(gout, y, scale, gx, pad)
    ref = torch.nn.grad.conv2d_input(x.shape, y.double(), gout.double(), stride=1, padding=pad)
    assert _rel(gx, ref) < TOL_TC, (name, _rel(gx, ref))
    rms = ((gx.double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert rms < 4e-3, (name, rms)


def test_conv_autograd_matches_aten():
    from po2_quantization_b200 import ops
    ops.set_dgrad_mode("aten")                    # this test pins the ATen formula; dgrad-on-K3 is tested above
    ops.set_wgrad_mode("aten")
    torch.backends.cudnn.allow_tf32 = False       # both backward passes in true fp32
    x, y, codes, scale = _make(RESNET[3])
    x1 = x.clone().requires_grad_(True)
    w1 = y.clone().requires_grad_(True)
    out = torch.ops.po2.conv2d(x1, w1, scale, 1, 1, 1, 0)
    g = torch.randn_like(out)
    out.backward(g)
    x2 = x.clone().requires_grad_(True)
    w2 = y.clone().requires_grad_(True)
    F.conv2d(x2, w2, None, 1, 1).backward(g)
    ops.set_dgrad_mode("tc")
    assert torch.allclose(x1.grad, x2.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(w1.grad, w2.grad, rtol=1e-4, atol=1e-3)
    # and with the tensor-core data gradient: same weight gradient, data gradient within bf16 rounding
    x3 = x.clone().requires_grad_(True)
    w3 = y.clone().requires_grad_(True)
    torch.ops.po2.conv2d(x3, w3, scale, 1, 1, 1, 0).backward(g)
    assert torch.allclose(w3.grad, w2.grad, rtol=1e-4, atol=1e-3)
    assert _rel(x3.grad, x2.grad.double()) < TOL_TC
    # and with the tensor-core weight gradient too: both operands rounded to bf16
    ops.set_wgrad_mode("tc")
    x4 = x.clone().requires_grad_(True)
    w4 = y.clone().requires_grad_(True)
    torch.ops.po2.conv2d(x4, w4, scale, 1, 1, 1, 0).backward(g)
    assert _rel(w4.grad, w2.grad.double()) < TOL_TC


WGRAD_CASES = [RESNET[0], RESNET[3], RESNET[6], MOBILENET[1], MOBILENET[4], ODD[1],
               ("wg 24->40 3x3 @12x20", 5, 24, 12, 20, 40, 3, 1, 1, 1), ("wg 128->128 3x3 @8", 8, 128, 8, 8, 128, 3, 1, 1, 1),
               ("wg 100->72 1x1 @7x9", 3, 100, 7, 9, 72, 1, 1, 0, 1), ("wg 16->16 3x3 @32 B=128", 128, 16, 32, 32, 16, 3, 1, 1, 1)]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: c[0])
def test_conv_wgrad_tensor_core_path(case):
    """Weight gradient on the tcgen05 kernel (x and grad rounded to bf16, fp32 accumulation over all
    pixels in TMEM, fixed-order partial sums) against fp64: rel 1e-2 like the forward, deterministic."""
    from po2_quantization_b200 import ops
    name, B, C, H, W, K, k, stride, pad, groups = case
    if stride != 1 or groups != 1 or K > 128:
        pytest.skip("shape not taken by the wgrad kernel")
    g0 = torch.Generator(device="cuda").manual_seed(B + C + K)
    x = torch.randn(B, C, H, W, device="cuda", generator=g0)
    gout = torch.randn(B, K, H, W, device="cuda", generator=g0)
    gw = torch.full((K, C, k, k), float("nan"), device="cuda")
    assert ops.conv2d_wgrad_out(gout, x, gw, pad, 0)
    ref = torch.nn.grad.conv2d_weight(x.double(), (K, C, k, k), gout.double(), stride=1, padding=pad)
    assert _rel(gw, ref) < TOL_TC, (name, _rel(gw, ref))
    rms = ((gw.double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert rms < 4e-3, (name, rms)
    gw2 = torch.empty_like(gw)
    assert ops.conv2d_wgrad_out(gout, x, gw2, pad)
    assert torch.equal(gw, gw2), "wgrad must be deterministic"
    # operands that are exact in bf16 make the result exact up to fp32 accumulation order
    xb, gb = x.bfloat16().float(), gout.bfloat16().float()
    assert ops.conv2d_wgrad_out(gb, xb, gw2, pad)
    ref_b = torch.nn.grad.conv2d_weight(xb.double(), (K, C, k, k), gb.double(), stride=1, padding=pad)
    assert _rel(gw2, ref_b) < 1e-5, (name, _rel(gw2, ref_b))


WGRAD_TMA_CASES = [RESNET[0], RESNET[3], RESNET[6], MOBILENET[1],
                   ("wg 8->8 3x3 @32 B=3", 3, 8, 32, 32, 8, 3, 1, 1, 1), ("wg 16->24 3x3 @8 B=6", 6, 16, 8, 8, 24, 3, 1, 1, 1),
                   ("wg 24->40 3x3 @4x4", 9, 24, 4, 4, 40, 3, 1, 1, 1), ("wg 64->128 3x3 @8 (M = 128)", 8, 64, 8, 8, 128, 3, 1, 1, 1),
                   ("wg 40->72 1x1 @6x6", 3, 40, 6, 6, 72, 1, 1, 0, 1), ("wg 32->128 1x1 @56", 4, 32, 56, 56, 128, 1, 1, 0, 1),
                   ("wg 16->16 3x3 @16x32 (H != W)", 5, 16, 16, 32, 16, 3, 1, 1, 1)]


@pytest.mark.parametrize("case", WGRAD_TMA_CASES, ids=lambda c: c[0])
def test_conv_wgrad_tma_tf32_path(case):
    """K5T (csrc/po2_wgrad_tma.cuh): go and x staged by tensor-map TMA straight from fp32 NCHW, tf32 operands (the
    low 13 mantissa bits are ignored by the tensor core, as in the reference's own cuDNN TF32 kernels), the
    two column-shifted copies of x derived inside shared memory.  Against fp64: rel 2e-3; deterministic;
    exact up to accumulation order when the operands are tf32-exact."""
    from po2_quantization_b200 import ops, _lib
    name, B, C, H, W, K, k, stride, pad, groups = case
    assert _lib.load().po2_conv2d_wgrad_kernel_kind(B, C, H, W, K, k, k, stride, pad, groups, 2) == 2, name
    g0 = torch.Generator(device="cuda").manual_seed(B + C + K)
    x = torch.randn(B, C, H, W, device="cuda", generator=g0)
    gout = torch.randn(B, K, H, W, device="cuda", generator=g0)
    gw = torch.full((K, C, k, k), float("nan"), device="cuda")
    assert ops.conv2d_wgrad_out(gout, x, gw, pad, 2)
    ref = torch.nn.grad.conv2d_weight(x.double(), (K, C, k, k), gout.double(), stride=1, padding=pad)
    assert _rel(gw, ref) < TOL_TF32, (name, _rel(gw, ref))
    gw2 = torch.empty_like(gw)
    assert ops.conv2d_wgrad_out(gout, x, gw2, pad, 2)
    assert torch.equal(gw, gw2), "wgrad must be deterministic"
    trunc = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)       # tf32-exact operands
    xb, gb = trunc(x.clone()), trunc(gout.clone())
    assert ops.conv2d_wgrad_out(gb, xb, gw2, pad, 2)
    ref_b = torch.nn.grad.conv2d_weight(xb.double(), (K, C, k, k), gb.double(), stride=1, padding=pad)
    assert _rel(gw2, ref_b) < 1e-5, (name, _rel(gw2, ref_b))


BWD_CASES = [RESNET[1], RESNET[2], RESNET[4], RESNET[5], MOBILENET[9], MOBILENET[10],
             ("dw_144_s2_at8", 16, 144, 8, 8, 144, 3, 2, 1, 144), ("dw_24_s1_odd", 3, 24, 7, 9, 24, 3, 1, 1, 24)]


@pytest.mark.parametrize("case", BWD_CASES, ids=[c[0] for c in BWD_CASES])
def test_conv_backward_stride2_and_depthwise_leave_aten(case):
    """Autograd of models/quantized_conv.py:36 for the stride-2 layers of models/resnet.py (zero-inserted output
    gradient -> the stride-1 tcgen05 kernels) and the depthwise layers of models/mobilenet.py (CUDA-core kernels
    of csrc/po2_conv_bwd.cu): gradients against fp64 autograd of F.conv2d, and no aten.convolution_backward."""
    from po2_quantization_b200 import ops
    name, B, C, H, W, K, k, stride, pad, groups = case
    x, y, codes, scale = _make(case)
    x.requires_grad_(True)
    w = y.clone().requires_grad_(True)
    ops._noted.discard("conv_backward_aten")
    out = torch.ops.po2.conv2d(x, w, scale, stride, pad, groups, 2)
    g = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
    out.backward(g)
    assert "conv_backward_aten" not in ops._noted, "the backward fell back to aten.convolution_backward"
    xd, wd = x.detach().double().requires_grad_(True), y.double().requires_grad_(True)
    F.conv2d(xd, wd, None, stride, pad, 1, groups).backward(g.double())
    tol = TOL_FP32 * 10 if groups > 1 else TOL_TF32
    assert _rel(x.grad, xd.grad) < tol, (name, _rel(x.grad, xd.grad))
    assert _rel(w.grad, wd.grad) < tol, (name, _rel(w.grad, wd.grad))
    x.grad = None
    w2 = y.clone().requires_grad_(True)
    torch.ops.po2.conv2d(x, w2, scale, stride, pad, groups, 2).backward(g)
    assert torch.equal(w.grad, w2.grad), "the weight gradient must be deterministic"


@pytest.mark.parametrize("plus", [False, True])
def test_module_qat_forward_backward_vs_oracle(plus):
    """QuantizedConv2d (QAT mode) against the oracle module on CPU: forward within the bf16
    tolerance, weight gradient straight-through."""
    import po2_quantization_b200 as P
    from oracle.po2_oracle_torch import PO2, PO2_PLUS, QuantizedConv2dOracle
    torch.manual_seed(3)
    Q = P.PowerOfTwoPlusQuantizer if plus else P.PowerOfTwoQuantizer
    m = P.QuantizedConv2d(32, 32, 3, stride=1, padding=1, quantize_fn=Q, bits=4).cuda()
    o = QuantizedConv2dOracle(32, 32, 3, stride=1, padding=1, quantize_fn=PO2_PLUS if plus else PO2, bits=4)
    o.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    assert list(m.state_dict().keys()) == ["weight"]
    x = torch.randn(16, 32, 16, 16)
    xg = x.cuda().requires_grad_(True)
    xc = x.clone().requires_grad_(True)
    out = m(xg)
    ref = o(xc)
    assert _rel(out.cpu(), ref.double()) < TOL_TC
    g = torch.randn_like(ref)
    out.backward(g.cuda())
    ref.backward(g)
    # weight gradient (straight-through): x and grad rounded to bf16 on the tensor-core kernel
    assert _rel(m.weight.grad.cpu(), o.weight.grad.double()) < TOL_TC
    assert _rel(xg.grad.cpu(), xc.grad.double()) < TOL_TC      # data gradient: grad rounded to bf16, PO2 weights exact
    # ... and exactly ATen's when the tensor-core gradient kernels are switched off
    from po2_quantization_b200 import ops
    ops.set_wgrad_mode("aten")
    try:
        m.weight.grad = None
        m(xg).backward(g.cuda())
        assert torch.allclose(m.weight.grad.cpu(), o.weight.grad, rtol=2e-3, atol=2e-3)
    finally:
        ops.set_wgrad_mode("tc")
    # models/quantized_conv.py:40-45: (sum((Q(w) - w)^2), numel) against the oracle quantizer in fp64
    e1, n1 = m.get_quantization_error()
    wd = o.weight.detach()
    ref_e = torch.sum(((PO2_PLUS if plus else PO2).forward(None, wd, bits=4).double() - wd.double()) ** 2).item()
    assert n1 == o.weight.numel()
    assert abs(e1.item() - ref_e) <= 1e-6 * ref_e, (e1.item(), ref_e)
    # no quantizer: (0, numel)
    m_fp = P.QuantizedConv2d(32, 32, 3, quantize_fn=None).cuda()
    assert m_fp.get_quantization_error() == (0, m_fp.weight.numel())


def test_ptq_quantize_model_and_forward_resnet20_top1():
    """BASELINE.json configs[0]: ResNet-20 PO2+ 4-bit PTQ forward, batch 128.  Known-answer MSE from
    the unmodified reference; logits and top-1 against the oracle model on CPU."""
    import po2_quantization_b200 as P
    from oracle.po2_oracle_torch import PO2_PLUS, QuantizedConv2dOracle, quantize_model_ref
    from tests import golden_util as G
    from workloads import resnet_cifar
    ka = G.load("known_answers.npz")
    torch.manual_seed(8)
    ref_model = resnet_cifar(20, 10, None, 4, conv_cls=QuantizedConv2dOracle)
    torch.manual_seed(8)
    model = resnet_cifar(20, 10, None, 4)
    model.load_state_dict(ref_model.state_dict(), strict=True)
    model = model.cuda()
    mcopy = copy.deepcopy(model)                               # test.py:120 deep-copies before PTQ
    mse = P.quantize_model(mcopy, P.PowerOfTwoPlusQuantizer, 4)
    assert abs(mse - float(ka["resnet20_ptq_mse|po2+|4"])) <= 1e-5 * mse
    mse_ref = quantize_model_ref(ref_model, PO2_PLUS, 4)
    assert abs(mse - mse_ref) <= 1e-5 * mse
    for a, b in zip(mcopy.parameters(), ref_model.parameters()):
        assert torch.equal(a.detach().cpu(), b.detach()), "PTQ weights differ from the oracle's"
    mcopy.eval()
    ref_model.eval()
    x = torch.randn(128, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        from po2_quantization_b200 import ops
        ops.LAUNCHES = 0
        logits = mcopy(x.cuda()).cpu()
        assert ops.LAUNCHES >= 20, "the PTQ forward did not go through the po2 conv kernels"
        ref = ref_model(x)
        assert _rel(logits, ref.double()) < TOL_TC
        margin = ref.topk(2, dim=1).values
        decisive = (margin[:, 0] - margin[:, 1]) > 2e-2 * ref.abs().max()
        assert torch.equal(logits.argmax(1)[decisive], ref.argmax(1)[decisive])
        ops.set_conv_mode("fp32")
        try:
            logits32 = mcopy(x.cuda()).cpu()
        finally:
            ops.set_conv_mode(ops.DEFAULT_CONV_MODE)
        assert _rel(logits32, ref.double()) < 1e-4
        assert torch.equal(logits32.argmax(1), ref.argmax(1)), "top-1 differs in fp32-accumulate mode"
    # a second deepcopy keeps the tag; an in-place weight update invalidates it
    m2 = copy.deepcopy(mcopy)
    conv = next(m for m in m2.modules() if isinstance(m, P.QuantizedConv2d))
    assert conv._po2_ptq[0] == conv.weight._version
    with torch.no_grad():
        conv.weight.mul_(1.0)
    assert conv._po2_ptq[0] != conv.weight._version


def test_ptq_forward_mobilenetv2_matches_oracle():
    """BASELINE.json configs[2]: MobileNetV2 PO2+ 4-bit (17 depthwise + 33 pointwise quantized convs)."""
    import po2_quantization_b200 as P
    from oracle.po2_oracle_torch import PO2_PLUS, QuantizedConv2dOracle, quantize_model_ref
    from po2_quantization_b200 import ops
    from workloads import mobilenet_v2_cifar
    torch.manual_seed(8)
    ref_model = mobilenet_v2_cifar(10, None, 4, conv_cls=QuantizedConv2dOracle)
    model = mobilenet_v2_cifar(10, None, 4)
    model.load_state_dict(ref_model.state_dict(), strict=True)
    model = model.cuda()
    mse = P.quantize_model(model, P.PowerOfTwoPlusQuantizer, 4)
    mse_ref = quantize_model_ref(ref_model, PO2_PLUS, 4)
    assert abs(mse - mse_ref) <= 1e-5 * mse_ref
    model.eval(); ref_model.eval()
    x = torch.randn(32, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ops.LAUNCHES = 0
        logits = model(x.cuda()).cpu()
        assert ops.LAUNCHES >= 50
        ref = ref_model(x)
        assert _rel(logits, ref.double()) < TOL_TC
        ops.set_conv_mode("fp32")
        try:
            logits32 = model(x.cuda()).cpu()
        finally:
            ops.set_conv_mode(ops.DEFAULT_CONV_MODE)
        assert _rel(logits32, ref.double()) < 1e-4


@pytest.mark.parametrize("img,patch,batch", [((32, 32), (1, 1), 16), ((224, 224), (1, 1), 2), ((64, 64), (2, 2), 4)])
def test_ptq_forward_mobilevit_8bit_matches_oracle(img, patch, batch):
    """BASELINE.json configs[3]: MobileViT-xs PO2+ 8-bit PTQ inference (33 quantized convs; 8-bit
    levels reach 2^-127).  224x224 is built directly with patch_size (1,1) as SURVEY.md section 8d says."""
    import po2_quantization_b200 as P
    from oracle.po2_oracle_torch import PO2_PLUS, QuantizedConv2dOracle, quantize_model_ref
    from po2_quantization_b200 import ops
    from workloads import mobilevit_xs
    torch.manual_seed(8)
    ref_model = mobilevit_xs(img, 1000, patch, None, 8, conv_cls=QuantizedConv2dOracle)
    model = mobilevit_xs(img, 1000, patch, None, 8)
    model.load_state_dict(ref_model.state_dict(), strict=True)
    model = model.cuda()
    mse = P.quantize_model(model, P.PowerOfTwoPlusQuantizer, 8)
    mse_ref = quantize_model_ref(ref_model, PO2_PLUS, 8)
    assert abs(mse - mse_ref) <= 1e-5 * mse_ref
    for a, b in zip(model.parameters(), ref_model.parameters()):
        assert torch.equal(a.detach().cpu(), b.detach()), "PTQ weights differ from the oracle's"
    model.eval(); ref_model.eval()
    x = torch.randn(batch, 3, *img, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ops.LAUNCHES = 0
        logits = model(x.cuda()).cpu()
        assert ops.LAUNCHES >= 33
        ref = ref_model(x)
    assert _rel(logits, ref.double()) < TOL_TC


@pytest.mark.parametrize("case", [RESNET[0], RESNET[1], RESNET[3], RESNET[6], MOBILENET[1], MOBILENET[3], MOBILENET[9], ODD[5]], ids=lambda c: c[0])
@pytest.mark.parametrize("plus", [False, True])
def test_fused_qat_forward_op_equals_quantize_then_conv(case, plus):
    """po2::qconv2d (quantizer kernel emits the packed operand itself) == quantize, then conv:
    bitwise the same quantized weight, scale and output, and the same gradients."""
    name, B, C, H, W, K, k, stride, pad, groups = case
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(min(B, 16), C, H, W, device="cuda", generator=g, requires_grad=True)
    w = (torch.randn(K, C // groups, k, k, device="cuda", generator=g) * 0.1).requires_grad_(True)
    out, qw, scale = torch.ops.po2.qconv2d(x, w, 4, 1, plus, stride, pad, groups, 0)
    y2, s2 = torch.ops.po2.quantize_scaled(w, 4, 1, plus)
    out2 = torch.ops.po2.conv2d(x, y2, s2, stride, pad, groups, 0)
    assert torch.equal(qw, y2) and scale.item() == s2.item()
    assert torch.equal(out, out2), name
    go = torch.randn_like(out)
    gx1, gw1 = torch.autograd.grad(out, (x, w), go)
    gx2, gw2 = torch.autograd.grad(out2, (x, w), go)
    assert torch.allclose(gx1, gx2, rtol=1e-5, atol=1e-6)
    assert torch.allclose(gw1, gw2, rtol=1e-3, atol=1e-3)      # cuDNN wgrad accumulates with atomics


def test_fused_qat_forward_scale_is_never_read_early():
    """The conv behind the fused quantize+pack kernel starts under programmatic dependent launch; its
    epilogue must not read `scale` before the quantizer has written it.  Alternate weights whose
    scales differ by 2^20 so that a recycled scale buffer holding the previous value is a gross error."""
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(8, 64, 8, 8, device="cuda", generator=g)
    ws = [torch.randn(64, 64, 1, 1, device="cuda", generator=g) * f for f in (2.0 ** -10, 2.0 ** 10)]
    refs = []
    for w in ws:
        y, s = torch.ops.po2.quantize_scaled(w, 4, 1, False)
        refs.append(torch.ops.po2.conv2d(x, y, s, 1, 0, 1, 0))
    torch.cuda.synchronize()
    for it in range(400):
        out, _, scale = torch.ops.po2.qconv2d(x, ws[it & 1], 4, 1, False, 1, 0, 1, 0)
        assert torch.equal(out, refs[it & 1]), it
        del out, scale


def test_fused_qat_forward_nan_weight_propagates():
    x = torch.randn(2, 16, 8, 8, device="cuda")
    w = torch.randn(16, 16, 3, 3, device="cuda")
    w[3, 2, 1, 1] = float("nan")
    out, qw, scale = torch.ops.po2.qconv2d(x, w, 4, 1, False, 1, 1, 1, 0)
    assert torch.isnan(qw).all() and torch.isnan(out).all()      # the reference: any NaN -> all NaN


def test_static_weight_pack_cache_single_launch():
    """PTQ-tagged module in no-grad mode: the packed operand is built once per (weight version, input
    shape) and every later forward is one launch with bitwise the same result."""
    import po2_quantization_b200 as P
    from po2_quantization_b200 import ops
    torch.manual_seed(1)
    seq = torch.nn.Sequential(P.QuantizedConv2d(32, 64, 3, 1, 1), P.QuantizedConv2d(64, 48, 1, 1, 0)).cuda()
    P.quantize_model(seq, P.PowerOfTwoPlusQuantizer, 4)
    x = torch.randn(8, 32, 16, 16, device="cuda")
    ref = seq[1](seq[0](x))                                   # grad mode on: regular path (pack + conv)
    with torch.no_grad():
        ops.LAUNCHES = 0
        a = seq(x)
        first = ops.LAUNCHES
        ops.LAUNCHES = 0
        b = seq(x)
        assert ops.LAUNCHES == 2 and first == 4               # 2 packs + 2 convs, then 2 convs only
        assert torch.equal(a, b) and torch.equal(a, ref.detach())
        c = seq(x[:4])                                        # new input shape -> new plan, re-packed
        assert torch.equal(c, ref.detach()[:4])
        seq[0].weight.mul_(2.0)                               # weight changed: tag invalid -> nn.Conv2d path
        d = seq(x)
        assert not torch.equal(d, a)


@pytest.mark.parametrize("compute", [0, 2], ids=["bf16", "tf32"])
def test_conv_tensor_core_path_random_shapes(compute):
    """Shape fuzz of the tcgen05 kernel: channel counts that need padding, K chunks, N tiles, strips
    spanning several images, widths that do / do not take the 128-bit producer path, both strides."""
    import random
    rnd = random.Random(1234)
    for it in range(60):
        k = rnd.choice([1, 3])
        stride = rnd.choice([1, 1, 2])
        B = rnd.choice([1, 2, 3, 7, 16])
        C = rnd.choice([3, 8, 16, 24, 40, 72, 100, 144, 200])
        K = rnd.choice([5, 16, 24, 40, 72, 136, 272, 300])
        H = rnd.choice([1, 2, 3, 4, 7, 8, 12, 16, 20, 33])
        W = rnd.choice([1, 2, 4, 5, 8, 12, 16, 28, 36])
        pad = 1 if k == 3 else 0
        if (H + 2 * pad - k) // stride + 1 <= 0 or (W + 2 * pad - k) // stride + 1 <= 0:
            continue
        g = torch.Generator(device="cuda").manual_seed(it)
        x = torch.randn(B, C, H, W, device="cuda", generator=g)
        w = torch.randn(K, C, k, k, device="cuda", generator=g) * 0.1
        y, codes, scale, _, _ = torch.ops.po2.quantize_full(w, 4, 1, bool(it & 1))
        tol = TOL_TC if compute == 0 else TOL_TF32
        out = torch.ops.po2.conv2d(x, y, scale, stride, pad, 1, compute)
        ref = _ref(x, y, stride, pad, 1)
        assert out.shape == ref.shape, (it, B, C, H, W, K, k, stride)
        err = _rel(out, ref)
        assert err < tol, (it, B, C, H, W, K, k, stride, err)
        if stride == 1:                                       # data gradient on the same kernel
            from po2_quantization_b200 import ops
            go = torch.randn_like(out)
            gx = torch.empty_like(x)
            if ops.conv2d_dgrad_out(go, y, scale, gx, pad, compute):
                gref = torch.nn.grad.conv2d_input(x.shape, y.double(), go.double(), stride=1, padding=pad)
                assert _rel(gx, gref) < tol, (it, "dgrad", B, C, H, W, K, k)


def test_depthwise_random_shapes():
    import random
    rnd = random.Random(99)
    for it in range(30):
        stride = rnd.choice([1, 2])
        B, C = rnd.choice([1, 3, 8]), rnd.choice([3, 16, 96, 130])
        H, W = rnd.choice([1, 2, 4, 7, 8, 16, 28]), rnd.choice([1, 2, 4, 6, 8, 16, 28])
        x = torch.randn(B, C, H, W, device="cuda")
        w = torch.randn(C, 1, 3, 3, device="cuda") * 0.2
        y, _, scale, _, _ = torch.ops.po2.quantize_full(w, 4, 1, True)
        out = torch.ops.po2.conv2d(x, y, scale, stride, 1, C, 0)
        ref = _ref(x, y, stride, 1, C)
        assert _rel(out, ref) < TOL_FP32, (it, B, C, H, W, stride)


@pytest.mark.parametrize("family", ["resnet20", "mobilenet", "mobilevit"])
def test_qat_training_step_matches_oracle_model(family):
    """One QAT forward+backward of each model family (BASELINE.json configs[1]-[3] in QAT mode) against
    the oracle model on CPU.  In fp32-accumulate mode every gradient must agree tightly (this pins
    the autograd plumbing: straight-through estimator, saved tensors, grouped / strided layers).  In
    tensor-core mode (bf16 activations) the loss and logits stay within the bf16 tolerance; early-layer
    gradients of a train-mode-BatchNorm network amplify ANY rounding (stock cuDNN TF32 -- the
    reference's own GPU default -- is already 5 % off the fp32 CPU gradient at the stem of ResNet-20
    with this batch; bf16 is 4x coarser), so there only direction and scale are checked."""
    import po2_quantization_b200 as P
    from oracle.po2_oracle_torch import PO2, QuantizedConv2dOracle
    from po2_quantization_b200 import ops
    from workloads import mobilenet_v2_cifar, mobilevit_xs, resnet_cifar
    torch.manual_seed(8)
    build = {"resnet20": lambda q, c: resnet_cifar(20, 10, q, 4, conv_cls=c),
             "mobilenet": lambda q, c: mobilenet_v2_cifar(10, q, 4, conv_cls=c),
             "mobilevit": lambda q, c: mobilevit_xs((32, 32), 10, (1, 1), q, 4, conv_cls=c)}[family]
    # BatchNorm in eval mode (running statistics): with 16 samples and 1x1 feature maps, train-mode BN
    # in MobileNet / MobileViT turns summation-order noise into O(1) logit changes on ANY backend, which
    # would make this a test of chaos, not of the kernels.  ResNet-20 keeps train-mode BN.
    bn_train = family == "resnet20"
    ref = build(PO2, QuantizedConv2dOracle).train(bn_train)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(16, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (16,), generator=g)
    crit = torch.nn.CrossEntropyLoss()
    lr = ref(x)
    loss_r = crit(lr, y)
    loss_r.backward()
    pr = dict(ref.named_parameters())
    old_tf32 = torch.backends.cudnn.allow_tf32
    try:
        for conv_mode, dgrad_mode, tol_logits, tol_grad in (("fp32", "aten", 5e-3, 3e-2), ("tc", "tc", 5e-2, None)):
            ops.set_conv_mode(conv_mode)
            ops.set_dgrad_mode(dgrad_mode)
            torch.backends.cudnn.allow_tf32 = False
            mine = build(P.PowerOfTwoQuantizer, None)
            mine.load_state_dict(ref.state_dict(), strict=True)
            mine = mine.cuda().train(bn_train)
            lm = mine(x.cuda())
            loss_m = crit(lm, y.cuda())
            loss_m.backward()
            assert abs(loss_m.item() - loss_r.item()) < tol_logits * max(1.0, abs(loss_r.item())), (conv_mode, loss_m.item(), loss_r.item())
            assert _rel(lm.detach().cpu(), lr.detach().double()) < tol_logits, conv_mode
            checked = 0
            gscale = max(v.grad.double().pow(2).mean().sqrt().item() for v in pr.values())
            for name, p in mine.named_parameters():
                assert p.grad is not None, f"{name} received no gradient (straight-through path broken?)"
                gm, gr = p.grad.cpu().double().flatten(), pr[name].grad.double().flatten()
                # a BN bias in front of another train-mode BN has an exactly-zero true gradient: both
                # sides then hold rounding noise only -- skip parameters whose gradient is negligible
                if p.grad.numel() < 64 or gr.pow(2).mean().sqrt().item() < 1e-7 * gscale:
                    continue
                rel_rms = ((gm - gr).pow(2).mean().sqrt() / (gr.pow(2).mean().sqrt() + 1e-12)).item()
                cos = (torch.dot(gm, gr) / (gm.norm() * gr.norm() + 1e-30)).item()
                if tol_grad is not None:
                    assert rel_rms < tol_grad, (family, conv_mode, name, rel_rms)
                elif family == "resnet20" and p.dim() >= 2:
                    # bf16 mode: direction and scale only, and only for the well-conditioned family (see docstring)
                    assert cos > 0.7 and 0.5 < (gm.norm() / gr.norm()).item() < 2.0, (family, conv_mode, name, cos, rel_rms)
                else:
                    assert torch.isfinite(gm).all(), (family, conv_mode, name)
                checked += 1
            assert checked >= 10, checked
    finally:
        ops.set_conv_mode(ops.DEFAULT_CONV_MODE)
        ops.set_dgrad_mode("tc")
        torch.backends.cudnn.allow_tf32 = old_tf32


def test_tf32_mode_end_to_end_module_paths():
    """conv mode 'tf32' through every module path: QAT op (fused quantize+pack), PTQ op, packed cache,
    data gradient -- all within the tf32 tolerance, and tighter than bf16 on the same inputs."""
    import po2_quantization_b200 as P
    from po2_quantization_b200 import ops
    torch.manual_seed(2)
    x = torch.randn(8, 32, 16, 16, device="cuda", requires_grad=True)
    conv = P.QuantizedConv2d(32, 64, 3, 1, 1, quantize_fn=P.PowerOfTwoPlusQuantizer, bits=4).cuda()
    qw, scale = torch.ops.po2.quantize_scaled(conv.weight.detach(), 4, 1, True)
    ref = _ref(x.detach(), qw, 1, 1, 1)
    go = torch.randn(8, 64, 16, 16, device="cuda")
    gref = torch.nn.grad.conv2d_input(x.shape, qw.double(), go.double(), stride=1, padding=1)
    errs = {}
    try:
        for mode in ("tc", "tf32"):
            ops.set_conv_mode(mode)
            x.grad = None
            out = conv(x)
            out.backward(go)
            errs[mode] = (_rel(out.detach(), ref), _rel(x.grad, gref))
        assert errs["tf32"][0] < TOL_TF32 and errs["tf32"][1] < TOL_TF32
        assert errs["tf32"][0] < errs["tc"][0] and errs["tf32"][1] < errs["tc"][1]
        # PTQ + packed cache in tf32
        seq = torch.nn.Sequential(P.QuantizedConv2d(32, 48, 1, 1, 0)).cuda()
        P.quantize_model(seq, P.PowerOfTwoQuantizer, 4)
        with torch.no_grad():
            a = seq(x.detach())
            b = seq(x.detach())
        r2 = _ref(x.detach(), seq[0].weight.detach(), 1, 0, 1)
        assert torch.equal(a, b) and _rel(a, r2) < TOL_TF32
    finally:
        ops.set_conv_mode(ops.DEFAULT_CONV_MODE)


def test_weight_prefetch_equals_inline_quantization():
    """enable_weight_prefetch: all weights quantized by one multi-tensor launch before the forward.
    Same arithmetic as the inline path, so with deterministic cuDNN (the stem conv) the models stay
    bitwise equal -- eagerly over optimizer steps (weights change, the prefetch must notice) and
    under CUDA-graph replay."""
    import po2_quantization_b200 as P
    from workloads import resnet_cifar
    torch.manual_seed(12)
    torch.backends.cudnn.deterministic = True
    a = resnet_cifar(20, 10, P.PowerOfTwoQuantizer, 4).cuda().train()
    b = copy.deepcopy(a)
    P.enable_weight_prefetch(b)
    oa = torch.optim.SGD(a.parameters(), lr=0.05, momentum=0.9)
    ob = torch.optim.SGD(b.parameters(), lr=0.05, momentum=0.9)
    x = torch.randn(32, 3, 32, 32, device="cuda")
    t = torch.randint(0, 10, (32,), device="cuda")
    for it in range(3):                                   # step 0 records shapes, steps 1-2 run prefetched
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad()
            F.cross_entropy(m(x), t).backward()
            o.step()
        used = sum(1 for m in b.modules() if getattr(m, "_po2_prefetch", None))
        assert (used > 0) == (it > 0), (it, used)
        for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
            assert torch.equal(pa, pb), (it, n)
    # CUDA-graph capture of a whole step with the side-stream fork inside
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        xs = x.clone()

        def step(m, o):
            o.zero_grad()
            F.cross_entropy(m(xs), t).backward()
            o.step()
        step(b, ob), step(a, oa)
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            step(b, ob)
        for _ in range(2):
            g.replay()
            step(a, oa)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    torch.backends.cudnn.deterministic = False
    for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
        assert torch.equal(pa, pb), n


EP_CASES = [RESNET[0], RESNET[1], RESNET[2], RESNET[3], RESNET[6], MOBILENET[1], MOBILENET[4], MOBILENET[7], MOBILENET[9],
            MOBILENET[10], MOBILENET[12], ODD[1], ODD[3], ODD[5]]


@pytest.mark.parametrize("act", [0, 1, 2, 3], ids=["none", "relu", "relu6", "silu"])
@pytest.mark.parametrize("case", EP_CASES, ids=[c[0] for c in EP_CASES])
def test_conv_epilogue_affine_residual_activation(case, act):
    """Inference epilogue (eval-mode BatchNorm folded into the conv, models/resnet.py:55-71): every kernel kind
    computes act(conv * a[k] + b[k] + residual) -- against fp64, with the conv's own operand tolerance."""
    from po2_quantization_b200 import ops
    name, B, C, H, W, K, k, stride, pad, groups = case
    x, y, codes, scale = _make(case)
    g0 = torch.Generator(device="cuda").manual_seed(K + act)
    a = torch.rand(K, device="cuda", generator=g0) + 0.5
    b = torch.randn(K, device="cuda", generator=g0)
    conv = _ref(x, y, stride, pad, groups)
    res = torch.randn(conv.shape, device="cuda", generator=g0) if act != 2 else None
    ref = conv * a.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1)
    if res is not None:
        ref = ref + res.double()
    ref = [lambda t: t, torch.relu, lambda t: t.clamp(0, 6), F.silu][act](ref)
    out = torch.ops.po2.conv2d_ep(x, y, scale, stride, pad, groups, 2, a, b, res, act)
    tol = TOL_FP32 * 10 if groups > 1 or k == 5 else TOL_TF32
    assert _rel(out, ref) < tol, (name, _rel(out, ref))
    packed = ops.conv2d_pack(y, scale, x.shape, stride, pad, groups, 2)
    if packed is not None:
        out2 = torch.ops.po2.conv2d_packed_ep(x, packed, scale, K, k, k, stride, pad, groups, 2, a, b, res, act)
        assert torch.equal(out, out2), name


@pytest.mark.parametrize("shape", [(128, 3, 32, 32, 16), (5, 3, 32, 32, 16), (300, 1, 28, 28, 8), (4, 4, 17, 23, 32), (2, 3, 40, 40, 24)],
                         ids=lambda s: "x".join(map(str, s)))
def test_stem_conv_weight_gradient_kernel(shape):
    """StemConv2d (the full-precision 3 -> 16 stem, models/resnet.py:99-102): forward is F.conv2d, the weight gradient
    runs on po2_conv2d_stem_wgrad -- against fp64 autograd, deterministic, state_dict and initialisation untouched;
    accelerate_stem re-classes a plain nn.Conv2d in place."""
    import po2_quantization_b200 as P
    from po2_quantization_b200 import ops
    B, C, H, W, K = shape
    torch.manual_seed(sum(shape))
    ref = torch.nn.Conv2d(C, K, 3, 1, 1, bias=False).cuda()
    m = torch.nn.Conv2d(C, K, 3, 1, 1, bias=False).cuda()
    m.load_state_dict(ref.state_dict())
    assert P.accelerate_stem(m) == 1 and isinstance(m, P.StemConv2d) and list(m.state_dict()) == ["weight"]
    x = torch.randn(B, C, H, W, device="cuda")
    go = torch.randn(B, K, H, W, device="cuda")
    before = ops.LAUNCHES
    y = m(x)
    y.backward(go)
    assert ops.LAUNCHES - before == 2                      # stem_wgrad_kernel + conv_wgrad_reduce_kernel
    first = m.weight.grad.clone()
    m.weight.grad = None
    m(x).backward(go)
    assert torch.equal(first, m.weight.grad)               # fixed-order partial sums
    w64 = ref.weight.detach().double().requires_grad_(True)
    torch.nn.functional.conv2d(x.double(), w64, None, 1, 1).backward(go.double())
    err = ((m.weight.grad.double() - w64.grad).abs().max() / w64.grad.abs().max()).item()
    assert err < 2e-5, err
    # an input that needs its own gradient goes to the library path with both gradients
    x2 = x.clone().requires_grad_(True)
    m.weight.grad = None
    m(x2).backward(go)
    assert x2.grad is not None and torch.allclose(m.weight.grad, first, rtol=2e-2, atol=2e-2 * first.abs().max().item())
