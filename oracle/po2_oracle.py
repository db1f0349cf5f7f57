"""TEST INFRASTRUCTURE ONLY -- numpy oracle for the PO2 / PO2+ quantizers.

An op-by-op CPU restatement of the reference's arithmetic, in numpy, used as the *checker* for the
sm_100a kernels.  Each function cites the reference lines it follows (paths are relative to the
reference checkout, ``mschoenb97/po2_quantization``).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference in the
build container, runs it on seeded inputs and commits the input/output vectors under
``tests/golden/``; ``tests/test_oracle_golden.py`` holds this oracle bit-exact against all of them
(fp32, bf16, fp16; bits 2..8; both quantizers; zeros, NaN, Inf, subnormal scale, exact-tie inputs).

The one libm-dependent step is ``torch.log2``.  This oracle uses a *correctly rounded* float log2
(float64 log2 rounded once to float32).  In the build container that agrees with torch's CPU
kernel on every rounding boundary of both quantizers (``tools/scan_boundaries.py --check``), so
"oracle == reference on CPU" holds bit-for-bit; where torch's CUDA ``log2f`` places a boundary
differently, that is enumerated in DESIGN.md rather than hidden.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "to_bf16", "bf16_to_f32", "f32_to_bf16_bits", "log2_cr_f32", "quantize", "po2", "po2_plus",
    "exponent_codes", "pack_codes", "unpack_codes", "dequantize_codes", "ste_backward",
    "quantization_sse",
]


# --------------------------------------------------------------------------------------------
# storage-dtype helpers (bf16 has no numpy dtype: carry it as uint16 bit patterns)
# --------------------------------------------------------------------------------------------
def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit pattern, round-to-nearest-even, NaN kept quiet (what torch does)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    rounding_bias = ((u >> 16) & 1) + np.uint32(0x7FFF)
    out = ((u + rounding_bias) >> 16).astype(np.uint16)
    nan = np.isnan(a)
    if np.any(nan):
        out = np.where(nan, np.uint16(0x7FC0), out)
    return out


def bf16_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def to_bf16(a: np.ndarray) -> np.ndarray:
    """Round float32 values to the bf16 grid, returned as float32."""
    return bf16_to_f32(f32_to_bf16_bits(a))


def _round_storage(a: np.ndarray, dtype: str) -> np.ndarray:
    """Round an fp32 intermediate to the storage dtype (torch rounds after *every* op)."""
    if dtype == "f32":
        return a.astype(np.float32)
    if dtype == "bf16":
        return to_bf16(a)
    if dtype == "f16":
        with np.errstate(over="ignore"):
            return a.astype(np.float16).astype(np.float32)
    raise ValueError(dtype)


def log2_cr_f32(v: np.ndarray) -> np.ndarray:
    """Correctly rounded float32 log2 (float64 evaluation, single final rounding)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log2(v.astype(np.float64)).astype(np.float32)


# --------------------------------------------------------------------------------------------
# the quantizers
# --------------------------------------------------------------------------------------------
def quantize(x: np.ndarray, bits: int = 4, fsr: int = 1, plus: bool = False, dtype: str = "f32",
             div15: str = "divide", return_parts: bool = False):
    """PO2 (plus=False) / PO2+ (plus=True) forward.

    ``x`` holds the input *values* as float32 (already on the storage grid when dtype != f32).
    Follows utils/quantizers.py:22-32 (PO2) and :42-52 (PO2+) one torch op per line; every
    intermediate is rounded to ``dtype`` because torch computes half types in fp32 and rounds
    after each op.

    div15: how ``abs_normalized_input / 1.5`` (utils/quantizers.py:47) is evaluated --
      "divide": IEEE division (torch CPU); "mulinv": ``v * (1/1.5f)`` (what torch's CUDA div
      kernel does for a Python-scalar divisor).
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    R = lambda a: _round_storage(a, dtype)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore", under="ignore"):
        # utils/quantizers.py:22  sign = torch.sign(input)   (torch.sign(nan) == 0)
        sign = (x > 0).astype(np.float32) - (x < 0).astype(np.float32)
        # :23  scale = torch.max(torch.abs(input))            (NaN propagates)
        scale = np.max(np.abs(x)) if x.size else np.float32(np.nan)
        scale = np.float32(scale)
        # :24-25  abs(input / scale)                          (IEEE division by a 0-dim tensor)
        v = np.abs(R(x / scale))
        # :27 / :47  log2 (+ the PO2+ shift), each op rounded to the storage dtype
        if plus:
            if div15 == "divide":
                w = R(v / np.float32(1.5))
            else:
                w = R(v * (np.float32(1.0) / np.float32(1.5)))
            l = R(R(log2_cr_f32(w)) + np.float32(0.5))
        else:
            l = R(log2_cr_f32(v))
        # :26-30 / :46-50  round half-to-even, clamp to [fsr - 2^(bits-1), fsr - 1]
        q = np.clip(np.rint(l), np.float32(fsr - 2 ** (bits - 1)), np.float32(fsr - 1))
        # :31 / :51  2 ** q   (exact, including subnormal 2^-127 .. 2^-149)
        qi = np.where(np.isnan(q), 0, q).astype(np.int32)
        p = np.ldexp(np.float32(1.0), qi).astype(np.float32)
        p = R(np.where(np.isnan(q), np.float32(np.nan), p))
        # :32 / :52  log_quant * sign * scale   (left to right)
        y = R(R(p * sign) * scale)
    if return_parts:
        return y, q, sign, scale
    return y


def po2(x, bits=4, fsr=1, dtype="f32"):
    """PowerOfTwoQuantizer.forward -- utils/quantizers.py:21-32."""
    return quantize(x, bits, fsr, False, dtype)


def po2_plus(x, bits=4, fsr=1, dtype="f32", div15="divide"):
    """PowerOfTwoPlusQuantizer.forward -- utils/quantizers.py:41-52."""
    return quantize(x, bits, fsr, True, dtype, div15)


def ste_backward(grad_output: np.ndarray) -> np.ndarray:
    """Straight-through estimator -- utils/quantizers.py:34-36, 54-56: grad_input = grad_output."""
    return grad_output


def quantization_sse(x: np.ndarray, y: np.ndarray) -> np.float32:
    """sum((q(w) - w)^2) -- models/quantized_conv.py:43 / utils/quantizers.py:149 (fp32 sum;
    summation order is torch's, so callers compare with a tolerance, not bitwise)."""
    d = (y.astype(np.float32) - x.astype(np.float32))
    return np.float32(np.sum(d * d, dtype=np.float32))


# --------------------------------------------------------------------------------------------
# sign+exponent codes (new in this build; defined from the reference's q and sign)
# --------------------------------------------------------------------------------------------
def exponent_codes(q: np.ndarray, sign: np.ndarray, bits: int, fsr: int = 1) -> np.ndarray:
    """code = signbit << (bits-1) | ((fsr-1) - q): magnitude field 0 is the largest level.
    An input of exactly 0 has sign 0 in the reference (utils/quantizers.py:22) and no code of its
    own: it is emitted as (+, clamp-min) and counted in the kernel's zero counter."""
    mag = (np.float32(fsr - 1) - q).astype(np.int64)
    return ((sign < 0).astype(np.int64) << (bits - 1) | mag).astype(np.uint8)


def code_container_bits(bits: int) -> int:
    return 4 if bits <= 4 else 8


def pack_codes(codes: np.ndarray, bits: int) -> np.ndarray:
    """bits<=4: two codes per byte, element 2i in the low nibble; else one code per byte."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8).ravel()
    if code_container_bits(bits) == 8:
        return codes.copy()
    if codes.size % 2:
        codes = np.concatenate([codes, np.zeros(1, np.uint8)])
    return (codes[0::2] | (codes[1::2] << 4)).astype(np.uint8)


def unpack_codes(packed: np.ndarray, n: int, bits: int) -> np.ndarray:
    packed = np.ascontiguousarray(packed, dtype=np.uint8).ravel()
    if code_container_bits(bits) == 8:
        return packed[:n].copy()
    out = np.empty(packed.size * 2, np.uint8)
    out[0::2] = packed & 0xF
    out[1::2] = packed >> 4
    return out[:n]


def dequantize_codes(codes: np.ndarray, scale, bits: int, fsr: int = 1, dtype: str = "f32"):
    """±2^q * scale from unpacked codes, with the reference's rounding (utils/quantizers.py:31-32)."""
    codes = codes.astype(np.int64)
    neg = (codes >> (bits - 1)) & 1
    mag = codes & ((1 << (bits - 1)) - 1)
    q = (fsr - 1) - mag
    p = np.ldexp(np.float32(1.0), q.astype(np.int32)).astype(np.float32)
    sgn = np.where(neg == 1, np.float32(-1), np.float32(1))
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        return _round_storage(_round_storage(p * sgn, dtype) * np.float32(scale), dtype)
