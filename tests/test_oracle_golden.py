"""The numpy oracle is held bit-exact against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import po2_oracle as O
from tests import golden_util as G

CASES = list(G.quantizer_cases())


@pytest.mark.parametrize("key", [c[0] for c in CASES])
def test_oracle_matches_reference_bits(key):
    _, name, dt, qn, bits, xb, yb = next(c for c in CASES if c[0] == key)
    fsr = 1
    if name.startswith("fsr"):
        fsr = int(name[3:])
    x = G.bits_to_f32(xb, dt)
    y = O.quantize(x, bits=bits, fsr=fsr, plus=(qn == "po2+"), dtype=dt)
    got = G.f32_to_bits(y, dt)
    ref_nan = G.nan_mask(yb, dt)
    got_nan = np.isnan(y)
    assert np.array_equal(ref_nan, got_nan), key
    ok = (got == yb) | ref_nan
    bad = np.flatnonzero(~ok)
    assert bad.size == 0, (key, bad[:5], x[bad[:5]], y[bad[:5]], G.bits_to_f32(yb, dt)[bad[:5]])


def test_known_answer_mse_change():
    """SURVEY.md section 4 / analysis.ipynb cell 14: po2 -> po2+ improves MSE by ~8.4 % at 4 bits."""
    ka = G.load("known_answers.npz")
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1 << 18).astype(np.float32)
    e = {}
    for plus in (False, True):
        y = O.quantize(x, 4, 1, plus)
        e[plus] = float(np.mean((y.astype(np.float64) - x) ** 2))
    change = (e[True] - e[False]) / e[False]
    ref_change = (ka["mse|po2+|4"] - ka["mse|po2|4"]) / ka["mse|po2|4"]
    assert abs(ref_change - (-0.084)) < 0.002
    assert abs(change - ref_change) < 0.01


def test_codes_roundtrip():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(1001).astype(np.float32)
    for bits in (2, 3, 4, 5, 8):
        for plus in (False, True):
            y, q, sign, scale = O.quantize(x, bits, 1, plus, return_parts=True)
            codes = O.exponent_codes(q, sign, bits)
            packed = O.pack_codes(codes, bits)
            assert packed.size == ((x.size + 1) // 2 if bits <= 4 else x.size)
            back = O.unpack_codes(packed, x.size, bits)
            assert np.array_equal(back, codes)
            y2 = O.dequantize_codes(back, scale, bits)
            assert np.array_equal(y2.view(np.uint32), y.view(np.uint32))


def test_levels_are_powers_of_two():
    rng = np.random.default_rng(2)
    x = rng.standard_normal(4096).astype(np.float32)
    for bits in (2, 3, 4):
        y = O.po2(x, bits)
        s = np.max(np.abs(x))
        lv = np.unique(np.abs(y) / s)
        assert lv.size <= 2 ** (bits - 1)
        assert np.all(np.log2(lv) == np.rint(np.log2(lv)))


TD = None


@pytest.mark.parametrize("key", [c[0] for c in CASES])
def test_torch_restatement_matches_reference_bits(key):
    """oracle/po2_oracle_torch.py (the CPU-baseline arm) == unmodified reference, bit for bit."""
    import torch
    from oracle.po2_oracle_torch import quantize_ref
    td = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}
    _, name, dt, qn, bits, xb, yb = next(c for c in CASES if c[0] == key)
    fsr = int(name[3:]) if name.startswith("fsr") else 1
    if dt == "f32":
        x = torch.from_numpy(xb.view(np.int32).copy()).view(torch.float32)
    else:
        x = torch.from_numpy(xb.view(np.int16).copy()).view(td[dt])
    y = quantize_ref(x, bits, fsr, qn == "po2+")
    got = y.view(torch.int32 if dt == "f32" else torch.int16).numpy().view(np.uint32 if dt == "f32" else np.uint16)
    ref_nan = G.nan_mask(yb, dt)
    assert np.array_equal(ref_nan, G.nan_mask(got, dt))
    assert np.all((got.ravel() == np.asarray(yb).ravel()) | ref_nan.ravel()), key
