"""Parity of the sm_100a quantizer kernels (through the C ABI / torch ops) against the oracle and
the committed reference vectors.  Bit-exact: integer comparison of the storage bit patterns."""
import numpy as np
import pytest
import torch

from oracle import po2_oracle as O
from tests import golden_util as G

pytestmark = pytest.mark.gpu

TD = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}


def _to_dev(bits_arr, dt):
    if dt == "f32":
        return torch.from_numpy(bits_arr.view(np.int32).copy()).view(torch.float32).cuda()
    return torch.from_numpy(bits_arr.view(np.int16).copy()).view(TD[dt]).cuda()


def _bits(t):
    t = t.detach().contiguous().cpu()
    if t.dtype == torch.float32:
        return t.view(torch.int32).numpy().view(np.uint32)
    return t.view(torch.int16).numpy().view(np.uint16)


def _assert_same(got_bits, ref_bits, dt, what):
    got_bits = np.asarray(got_bits).ravel()
    ref_bits = np.asarray(ref_bits).ravel()
    ref_nan = G.nan_mask(ref_bits, dt)
    got_nan = G.nan_mask(got_bits, dt)
    assert np.array_equal(ref_nan, got_nan), f"{what}: NaN pattern differs"
    bad = np.flatnonzero((got_bits != ref_bits) & ~ref_nan)
    assert bad.size == 0, (what, bad[:8], got_bits[bad[:8]], ref_bits[bad[:8]])


CASES = list(G.quantizer_cases())


@pytest.fixture(scope="module")
def P():
    import po2_quantization_b200 as P
    P.set_log2_flavor("ieee")
    return P


@pytest.mark.parametrize("key", [c[0] for c in CASES])
def test_golden_vectors_bit_exact(P, key):
    """CUDA kernels == unmodified reference (CPU) on every committed vector."""
    _, name, dt, qn, bits, xb, yb = next(c for c in CASES if c[0] == key)
    fsr = int(name[3:]) if name.startswith("fsr") else 1
    x = _to_dev(xb, dt).reshape(-1)
    Q = P.PowerOfTwoPlusQuantizer if qn == "po2+" else P.PowerOfTwoQuantizer
    y = Q.forward(None, x, bits=bits, fsr=fsr)
    assert y.dtype == x.dtype and y.shape == x.shape and y.data_ptr() != x.data_ptr()
    _assert_same(_bits(y), yb, dt, key)


@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("plus", [False, True])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 1023, 4099, 36864, 65536 + 5, (1 << 20) + 3])
def test_vs_oracle_sizes(P, dt, plus, n):
    rng = np.random.default_rng(n * 7 + plus)
    x32 = (rng.standard_normal(n) * rng.choice([1e-3, 0.05, 1.0, 37.0])).astype(np.float32)
    if n > 16:
        x32[5] = 0.0
        x32[11] = -0.0
    x32 = O._round_storage(x32, dt)
    xd = torch.from_numpy(x32).cuda().to(TD[dt])
    for bits in (2, 3, 4, 5, 8):
        y_ref, q, sign, scale = O.quantize(x32, bits, 1, plus, dt, return_parts=True)
        y, codes, s, zc, sse = torch.ops.po2.quantize_full(xd, bits, 1, plus)
        _assert_same(_bits(y), G.f32_to_bits(y_ref, dt), dt, f"y n={n} bits={bits}")
        assert s.item() == float(scale)
        ref_codes = O.pack_codes(O.exponent_codes(q, sign, bits), bits)
        assert np.array_equal(codes.cpu().numpy(), ref_codes), f"codes n={n} bits={bits}"
        assert zc.item() == int(np.sum(x32 == 0))
        ref_sse = float(np.sum((y_ref.astype(np.float64) - x32.astype(np.float64)) ** 2))
        assert abs(sse.item() - ref_sse) <= 1e-4 * max(ref_sse, 1e-30) + 1e-30
        # codes -> values round trip reproduces y (zeros have no code: mask them)
        back = torch.ops.po2.dequantize(codes, s, n, bits, 1, TD[dt])
        nz = torch.from_numpy(x32 != 0).cuda()
        assert torch.equal(back[nz].view(torch.int16 if dt != "f32" else torch.int32),
                           y[nz].view(torch.int16 if dt != "f32" else torch.int32))


@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_unaligned_views_and_noncontiguous(P, dt):
    rng = np.random.default_rng(3)
    base = O._round_storage(rng.standard_normal(5000).astype(np.float32), dt)
    bd = torch.from_numpy(base).cuda().to(TD[dt])
    for off in (1, 2, 3, 5):
        xv = bd[off:off + 4001]
        y = P.PowerOfTwoQuantizer.forward(None, xv, bits=4)
        ref = O.po2(base[off:off + 4001], 4, 1, dt)
        _assert_same(_bits(y), G.f32_to_bits(ref, dt), dt, f"offset {off}")
    # raw C-ABI call on a misaligned pointer takes the scalar kernel
    from po2_quantization_b200 import ops
    xv = bd[1:4002]
    y = torch.empty(4008, dtype=TD[dt], device="cuda")[1:4002]
    s = torch.empty((), dtype=torch.float32, device="cuda")
    codes = torch.zeros(2001, dtype=torch.uint8, device="cuda")
    ops.absmax_out(xv, s)
    ops.quantize_out(xv, y, s, 4, 1, True, codes=codes)
    yr, q, sg, sc = O.quantize(base[1:4002], 4, 1, True, dt, return_parts=True)
    _assert_same(_bits(y), G.f32_to_bits(yr, dt), dt, "scalar path")
    assert np.array_equal(codes.cpu().numpy(), O.pack_codes(O.exponent_codes(q, sg, 4), 4))
    x2 = bd[:4096].reshape(64, 64).t()          # non-contiguous input
    y2 = P.PowerOfTwoPlusQuantizer.forward(None, x2, bits=3)
    ref2 = O.po2_plus(base[:4096].reshape(64, 64).T.copy().ravel(), 3, 1, dt).reshape(64, 64)
    _assert_same(_bits(y2).ravel(), G.f32_to_bits(ref2.ravel(), dt), dt, "non-contiguous")


@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_two_pass_equals_fused_large(P, dt):
    """Streaming two-pass path (absmax + quantize, forward and reversed walk) == fused path == oracle."""
    from po2_quantization_b200 import ops
    n = (1 << 24) + 11
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(n, generator=g, device="cuda").to(TD[dt])
    s = torch.empty((), dtype=torch.float32, device="cuda")
    ops.absmax_out(x, s)
    assert s.item() == x.abs().max().item()
    outs = []
    for plus in (False, True):
        y1 = torch.empty_like(x)
        ops.quantize_out(x, y1, s, 4, 1, plus)
        y2 = torch.empty_like(x)
        s2 = torch.empty_like(s)
        ops.quantize_fused_out(x, y2, s2, 4, 1, plus)
        assert torch.equal(y1.view(torch.int16 if dt != "f32" else torch.int32),
                           y2.view(torch.int16 if dt != "f32" else torch.int32))
        assert s2.item() == s.item()
        outs.append(y1)
    # oracle on a slice that contains the max (so the scale agrees)
    imax = int(torch.argmax(x.abs()).item())
    lo = max(0, min(imax - 1000, n - 200000))
    xs = torch.cat([x[lo:lo + 200000], x[imax:imax + 1]]).float().cpu().numpy()
    for plus, y in zip((False, True), outs):
        ref = O.quantize(xs, 4, 1, plus, dt)
        got = torch.cat([y[lo:lo + 200000], y[imax:imax + 1]])
        _assert_same(_bits(got), G.f32_to_bits(ref, dt), dt, "large slice")


def test_full_size_properties(P):
    """BASELINE-size properties (2^28 fp32): idempotence, level set, sign, scale."""
    n = 1 << 28
    x = torch.randn(n, device="cuda")
    for Q, bits in ((P.PowerOfTwoQuantizer, 4), (P.PowerOfTwoPlusQuantizer, 4), (P.PowerOfTwoPlusQuantizer, 8)):
        y = Q.forward(None, x, bits=bits)
        s = x.abs().max()
        assert y.abs().max().item() == s.item()
        y2 = Q.forward(None, y, bits=bits)
        assert torch.equal(y2.view(torch.int32), y.view(torch.int32)), "not idempotent"
        assert torch.equal(torch.signbit(y), torch.signbit(x))
        nz = x != 0                                   # an exact 0.0 maps to 0 (sign() == 0)
        assert torch.equal(y[~nz], torch.zeros_like(y[~nz]))
        r = (y[nz].abs() / s)
        e = torch.log2(r)
        assert torch.equal(e, torch.round(e)), "levels are not powers of two times the scale"
        assert e.min().item() >= 1 - 2 ** (bits - 1) and e.max().item() == 0
        del y, y2, r, e


def test_autograd_ste(P):
    w = torch.randn(64, 16, 3, 3, device="cuda", requires_grad=True)
    g = torch.randn_like(w)
    for Q in (P.PowerOfTwoQuantizer, P.PowerOfTwoPlusQuantizer):
        w.grad = None
        y = Q.apply(w, 4)
        y.backward(g)
        assert torch.equal(w.grad, g)              # straight-through: utils/quantizers.py:34-36
    assert Q.backward(None, g)[0] is g
    gi = torch.zeros_like(g)
    torch.ops.po2.ste_backward(g, gi, False)
    assert torch.equal(gi, g)
    torch.ops.po2.ste_backward(g, gi, True)
    assert torch.equal(gi, g + g)
    gb = g.bfloat16()[:1001 * 3].contiguous()
    gi = torch.ones_like(gb)
    torch.ops.po2.ste_backward(gb, gi, True)
    assert torch.equal(gi, gb + 1)


def test_errors_are_loud(P):
    from po2_quantization_b200 import _lib
    with pytest.raises(Exception):
        P.PowerOfTwoQuantizer.forward(None, torch.randn(4), bits=4)       # CPU tensor: no fallback
    with pytest.raises(Exception):
        P.PowerOfTwoQuantizer.forward(None, torch.randn(4, device="cuda").double(), bits=4)
    with pytest.raises(_lib.Po2Error):
        P.PowerOfTwoQuantizer.forward(None, torch.randn(4, device="cuda"), bits=9)
    with pytest.raises(RuntimeError):
        P.PowerOfTwoQuantizer.forward(None, torch.empty(0, device="cuda"), bits=4)


def test_cuda_graph_capture(P):
    x = torch.randn(36864, device="cuda")
    y_eager = P.PowerOfTwoPlusQuantizer.forward(None, x, bits=4)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        P.PowerOfTwoPlusQuantizer.forward(None, x, bits=4)            # warm-up on the side stream
        torch.cuda.current_stream().synchronize()
        with torch.cuda.graph(g, stream=s):
            y = P.PowerOfTwoPlusQuantizer.forward(None, x, bits=4)
    x.copy_(torch.randn_like(x))
    g.replay()
    torch.cuda.synchronize()
    ref = O.po2_plus(x.cpu().numpy(), 4)
    assert np.array_equal(_bits(y), ref.view(np.uint32))
    assert y_eager.shape == y.shape


def _boundary_inputs(plus, scale, device):
    """fp32 values straddling every rounding boundary (same construction as make_golden.py)."""
    centre = 0x400000 if plus else 0x3504F3
    out = []
    for e in range(-126, 0):
        m = np.arange(centre - 40, centre + 41, dtype=np.int64)
        out.append((((e + 127) << 23) | m).astype(np.uint32))
    v = torch.from_numpy(np.concatenate(out).view(np.int32).copy()).view(torch.float32)
    s = torch.tensor(scale, dtype=torch.float32)
    x = v * s
    xb = x.view(torch.int32)
    x = torch.cat([torch.clamp(xb + d, min=0).view(torch.float32) for d in (-2, -1, 0, 1, 2)])
    x = x[x <= s]
    return torch.cat([x, s.reshape(1)]).to(device)


@pytest.mark.parametrize("plus", [False, True])
def test_torch_cuda_flavor_matches_stock_cuda_ops(P, plus):
    """With the torch_cuda boundary table the kernels reproduce the reference *run on this GPU*
    (stock ATen CUDA ops) bit for bit, including at every rounding boundary."""
    from oracle.po2_oracle_torch import quantize_ref
    Q = P.PowerOfTwoPlusQuantizer if plus else P.PowerOfTwoQuantizer
    P.set_log2_flavor("torch_cuda")
    try:
        g = torch.Generator(device="cuda").manual_seed(11)
        xs = [torch.randn(1 << 24, generator=g, device="cuda"),
              torch.randn(1 << 22, generator=g, device="cuda") ** 3 * 1e-3]
        xs += [_boundary_inputs(plus, sc, "cuda") for sc in (1.0, 1.337, 0.0517, 1.4142135, 1.3333334, 7.7e30)]
        for x in xs:
            for bits in (3, 4, 8):
                ours = Q.forward(None, x, bits=bits)
                stock = quantize_ref(x, bits, 1, plus)
                neq = (ours.view(torch.int32) != stock.view(torch.int32))
                assert int(neq.sum().item()) == 0, (bits, x[neq][:5], ours[neq][:5], stock[neq][:5])
    finally:
        P.set_log2_flavor("ieee")


def test_stock_torch_cuda_disagreements(P, tmp_path):
    """Enumerate (not assert) where stock torch CUDA ops -- the reference run on this GPU -- differ
    from the reference run on a CPU (== oracle == our default 'ieee' flavor)."""
    import json
    import os
    from oracle.po2_oracle_torch import quantize_ref
    rep = {}
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(1 << 26, generator=g, device="cuda")
    for plus in (False, True):
        for bits in (4, 8):
            Q = P.PowerOfTwoPlusQuantizer if plus else P.PowerOfTwoQuantizer
            ours = Q.forward(None, x, bits=bits)
            stock = quantize_ref(x, bits, 1, plus)
            rep[f"{'po2+' if plus else 'po2'}|{bits}|ieee_vs_stock_cuda_mismatch_of_2^26"] = int((ours != stock).sum().item())
    out = os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(rep, open(os.path.join(out, "stock_cuda_disagreements.json"), "w"), indent=1)
    print(rep)


@pytest.mark.parametrize("plus", [False, True])
@pytest.mark.parametrize("scale", [1.0, 1.337])
def test_exhaustive_fp32_patterns_vs_stock_cuda(P, plus, scale):
    """Every fp32 bit pattern with 0 <= |x| <= scale (about 1.07e9 of them, both quantizers, bits 2..8
    at scale 1, bits 4 and 8 at scale 1.337) against the reference run on this GPU (stock ATen CUDA ops), bitwise, under the
    torch_cuda boundary table.  The max is planted so that the tensor's scale is exactly `scale`."""
    from oracle.po2_oracle_torch import quantize_ref
    Q = P.PowerOfTwoPlusQuantizer if plus else P.PowerOfTwoQuantizer
    top = int(torch.tensor(scale, dtype=torch.float32).view(torch.int32).item())
    chunk = 1 << 27
    P.set_log2_flavor("torch_cuda")
    try:
        for start in range(0, top + 1, chunk):
            n = min(chunk, top + 1 - start)
            bits_ = torch.arange(start, start + n, device="cuda", dtype=torch.int64).to(torch.int32)
            x = bits_.view(torch.float32)
            # alternate signs, plant the maximum at the end
            x = torch.where((bits_ & 1).bool(), -x, x)
            x = torch.cat([x, torch.tensor([scale], device="cuda")])
            for nbits in ((2, 3, 4, 5, 6, 7, 8) if scale == 1.0 else (4, 8)):
                ours = Q.forward(None, x, bits=nbits)
                stock = quantize_ref(x, nbits, 1, plus)
                bad = ours.view(torch.int32) != stock.view(torch.int32)
                # NaN == NaN (only the all-zero chunk head: 0/scale is fine, so no NaNs expected)
                bad &= ~(torch.isnan(ours) & torch.isnan(stock))
                nbad = int(bad.sum().item())
                assert nbad == 0, (start, nbits, x[bad][:4], ours[bad][:4], stock[bad][:4])
            del x, bits_
    finally:
        P.set_log2_flavor("ieee")


def test_ieee_vs_torch_cuda_flavor_difference_is_the_enumerated_boundaries(P):
    """The two flavors may differ ONLY at the boundaries DESIGN.md enumerates: for po2 (4-bit range)
    not at all; for po2+ exactly at the three 1-ulp-shifted boundaries k = -6, -5, -4 (the k = -7
    boundary lies below the 4-bit clamp: both of its sides map to the minimum level)."""
    top = 0x3F800000
    chunk = 1 << 27
    diffs = {False: [], True: []}
    for plus in (False, True):
        Q = P.PowerOfTwoPlusQuantizer if plus else P.PowerOfTwoQuantizer
        for start in range(0, top + 1, chunk):
            n = min(chunk, top + 1 - start)
            bits_ = torch.arange(start, start + n, device="cuda", dtype=torch.int64).to(torch.int32)
            x = torch.cat([bits_.view(torch.float32), torch.ones(1, device="cuda")])
            P.set_log2_flavor("ieee")
            a = Q.forward(None, x, bits=4)
            P.set_log2_flavor("torch_cuda")
            b = Q.forward(None, x, bits=4)
            P.set_log2_flavor("ieee")
            d = (a.view(torch.int32) != b.view(torch.int32)).nonzero().flatten()
            diffs[plus] += [int(v) for v in x[d].view(torch.int32).cpu()]
    assert diffs[False] == []
    # cuda boundaries are 1 ulp below the cpu ones at k=-6..-4: exactly those three patterns flip
    assert sorted(diffs[True]) == [0x3C3FFFFE, 0x3CC00002, 0x3D3FFFFE], [hex(v) for v in diffs[True]]


def _half_patterns(dt, scale_bits):
    """all 2^16 bit patterns of a half type whose magnitude does not exceed `scale_bits` (both signs,
    +-0, subnormals), the scale itself planted last so that max|x| is exactly it"""
    u = np.arange(1 << 16, dtype=np.uint32)
    keep = (u & 0x7FFF) <= scale_bits
    return np.concatenate([u[keep], np.array([scale_bits], dtype=np.uint32)]).astype(np.uint16)


@pytest.mark.parametrize("dt", ["bf16", "f16"])
@pytest.mark.parametrize("plus", [False, True])
def test_exhaustive_half_patterns_both_flavors(P, dt, plus):
    """ALL 2^16 bit patterns of bf16 and fp16 as inputs (every intermediate of the
    reference is rounded to the storage dtype, utils/quantizers.py:22-32 / 42-52), bits 2..8, both
    quantizers, at several scales:
      * flavor "ieee"       bitwise against the numpy oracle (== the reference on CPU),
      * flavor "torch_cuda" bitwise against the reference's ops run on this GPU in the half dtype."""
    from oracle.po2_oracle_torch import quantize_ref
    Q = P.PowerOfTwoPlusQuantizer if plus else P.PowerOfTwoQuantizer
    one = 0x3F80 if dt == "bf16" else 0x3C00
    # 1.0, a non-power-of-two, the largest finite value, a tiny normal, a subnormal scale
    scales = [one, one + 0x2B, (0x7F7F if dt == "bf16" else 0x7BFF), (0x0100 if dt == "bf16" else 0x0500), 0x0007]
    for sb in scales:
        xb = _half_patterns(dt, sb)
        x = _to_dev(xb, dt)
        vals = G.bits_to_f32(xb, dt)
        for nbits in range(2, 9):
            P.set_log2_flavor("ieee")
            got = _bits(Q.forward(None, x, bits=nbits))
            ref = G.f32_to_bits(O.quantize(vals, nbits, 1, plus, dtype=dt), dt)
            _assert_same(got, ref, dt, f"ieee {dt} plus={plus} scale={sb:#x} bits={nbits}")
            P.set_log2_flavor("torch_cuda")
            try:
                got_c = _bits(Q.forward(None, x, bits=nbits))
            finally:
                P.set_log2_flavor("ieee")
            stock = _bits(quantize_ref(x, nbits, 1, plus))
            _assert_same(got_c, stock, dt, f"torch_cuda {dt} plus={plus} scale={sb:#x} bits={nbits}")


@pytest.mark.parametrize("plus", [False, True])
def test_flavor_difference_all_bits_is_the_table_difference(P, plus):
    """For every bit width 2..8 the two log2 flavors may differ ONLY between the boundaries where the
    compiled tables differ (16 po2 and 44 po2+ boundaries of 150), and never below
    the clamp: exhaustive over all fp32 magnitudes <= 1.  Together with the exhaustive torch_cuda check
    above this pins the default (ieee) flavor at every bit width, not only at 4 bits."""
    import json
    import os
    tdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "po2_quantization_b200", "tables")
    key = "f32|po2+" if plus else "f32|po2"
    ta = json.load(open(os.path.join(tdir, "oracle.json")))["tables"][key]
    tb = json.load(open(os.path.join(tdir, "torch_cuda.json")))["tables"][key]
    Q = P.PowerOfTwoPlusQuantizer if plus else P.PowerOfTwoQuantizer
    top = 0x3F800000
    chunk = 1 << 27
    ndiff = {b: 0 for b in range(2, 9)}
    lo_pat = {b: None for b in range(2, 9)}
    for start in range(0, top + 1, chunk):
        n = min(chunk, top + 1 - start)
        bits_ = torch.arange(start, start + n, device="cuda", dtype=torch.int64).to(torch.int32)
        x = torch.cat([bits_.view(torch.float32), torch.ones(1, device="cuda")])
        for nbits in range(2, 9):
            P.set_log2_flavor("ieee")
            a = Q.forward(None, x, bits=nbits)
            P.set_log2_flavor("torch_cuda")
            try:
                b = Q.forward(None, x, bits=nbits)
            finally:
                P.set_log2_flavor("ieee")
            d = (a.view(torch.int32) != b.view(torch.int32)).nonzero().flatten()
            ndiff[nbits] += int(d.numel())
            if d.numel():
                v = int(x[d].view(torch.int32).min().item())
                lo_pat[nbits] = v if lo_pat[nbits] is None else min(lo_pat[nbits], v)
    # expected: sum over boundaries k in (qmin, 0] of |threshold_ieee(k) - threshold_cuda(k)| patterns
    def thresholds(t):
        # tables/*.json: bounds_f32bits[i] = smallest bit pattern of v with raw(v) >= kmin + i
        return {t["kmin"] + i: int(v) for i, v in enumerate(t["bounds_f32bits"])}
    A, B = thresholds(ta), thresholds(tb)
    for nbits in range(2, 9):
        qmin = 1 - 2 ** (nbits - 1)
        expect = sum(abs(A[k] - B[k]) for k in A if qmin < k <= 0 and k in B)
        assert ndiff[nbits] == expect, (nbits, ndiff[nbits], expect)
    assert ndiff[2] == ndiff[3] == 0 and (ndiff[4] == (3 if plus else 0))


def test_lin_quantizers_cuda_kernel_matches_reference_vectors():
    """lin / lin+ (utils/quantizers.py:59-136) on the single-launch kernel: bit-exact against the vectors
    generated from the unmodified reference (tests/golden/lin_golden.npz)."""
    import po2_quantization_b200 as P
    from tests import golden_util as G
    z = G.load("lin_golden.npz")
    w = torch.from_numpy(z["w"]).cuda()
    launches = P.ops.LAUNCHES
    for name, cls in (("lin", P.LinearPowerOfTwoQuantizer), ("lin+", P.LinearPowerOfTwoPlusQuantizer)):
        for bits in (3, 4):
            got = cls.forward(None, w, bits=bits).cpu().numpy()
            assert np.array_equal(got.view(np.uint32), z[f"{name}|{bits}"].view(np.uint32)), (name, bits)
            g2 = cls.apply(w.clone().requires_grad_(True), bits)
            assert np.array_equal(g2.detach().cpu().numpy().view(np.uint32), z[f"{name}|{bits}"].view(np.uint32))
    assert P.ops.LAUNCHES == launches + 8            # one kernel launch per call


@pytest.mark.parametrize("shape", [(16, 16, 3, 3), (64, 32, 3, 3), (96, 16, 1, 1), (320, 960, 1, 1), (64, 128, 3, 3), (8, 3, 5, 5)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("plus", [False, True])
def test_lin_cuda_kernel_vs_op_by_op_form(shape, plus, monkeypatch):
    """The kernel against the reference's own op sequence run by torch on the CPU (the oracle for this
    quantizer).  The only licence: a channel whose <q,w>/<q,q> lands within rounding of a power-of-two
    tie may pick the neighbouring step (torch's fp32 summation order decides there), so per-channel
    disagreement must be rare and must be exactly a factor-of-two step."""
    import po2_quantization_b200 as P
    from po2_quantization_b200 import quantizers as Q
    g = torch.Generator().manual_seed(sum(shape) + int(plus))
    w = torch.randn(shape, generator=g) * 0.07
    cls = P.LinearPowerOfTwoPlusQuantizer if plus else P.LinearPowerOfTwoQuantizer
    for bits in (2, 4, 8):
        got = cls.forward(None, w.cuda(), bits=bits).cpu()
        monkeypatch.setenv("PO2_LIN", "aten")
        ref = Q._lin_forward(w, bits, 10, plus)              # CPU, op by op
        monkeypatch.setenv("PO2_LIN", "cuda")
        same = (got.view(torch.int32) == ref.view(torch.int32)).all(dim=3).all(dim=2).all(dim=0)   # per input channel
        assert same.float().mean().item() >= 0.97, (bits, same.float().mean().item())
        for c in (~same).nonzero().flatten().tolist():
            # both are k * 2^e grids; the steps differ by exactly one binade
            sg = got[:, c][got[:, c] != 0].abs().min().item()
            sr = ref[:, c][ref[:, c] != 0].abs().min().item()
            assert sg in (2 * sr, sr / 2, sr), (bits, c, sg, sr)


def test_lin_cuda_kernel_nan_and_constant_channels():
    import po2_quantization_b200 as P
    w = torch.randn(8, 4, 3, 3)
    w[:, 1] = 0.25                      # constant channel: step 0 -> 0/0 -> NaN in the reference too
    w[2, 2, 1, 1] = float("nan")        # NaN poisons its channel
    ref = P.quantizers._lin_forward(w, 4, 10, False)
    got = P.LinearPowerOfTwoQuantizer.forward(None, w.cuda(), bits=4).cpu()
    assert torch.equal(torch.isnan(got), torch.isnan(ref))
    ok = ~torch.isnan(ref)
    assert torch.equal(got[ok], ref[ok])
