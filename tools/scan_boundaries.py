"""Locate every rounding boundary of the reference's ``round(log2(.))`` pipelines by bisection.

The reference decides the exponent with a *float* log2 followed by round-half-even
(utils/quantizers.py:26-27 and :46-47), so the decision boundary in ``v = abs(x/scale)`` is a
property of the libm that evaluates it.  This tool treats the pipeline as a black box,

    po2 : raw(v) = round(log2(v))
    po2+: raw(v) = round(log2(v / 1.5) + 0.5)

evaluated by one of three backends -- ``torch_cpu`` (ATen CPU kernels, what the reference runs on
a CPU), ``torch_cuda`` (ATen CUDA kernels, what it runs on a GPU; run this on the B200 box) or
``oracle`` (oracle/po2_oracle.py, correctly rounded log2) -- and finds, for every integer k in
[-149, 0] and every storage dtype, the smallest v (as a bit pattern on the dtype's grid) with
raw(v) >= k.  It then checks that the pipeline is monotone in a +-W ulp window around each
boundary (so "one threshold per level" is a fact, not an assumption).

The resulting table *is* the definition the sm_100a kernel quantizes against (it never calls
libm); tables from different backends are diffed to enumerate the torch.log2 tie-boundary
disagreements.

    python tools/scan_boundaries.py --backend torch_cpu --out po2_quantization_b200/tables/torch_cpu.json
    python tools/scan_boundaries.py --diff a.json b.json
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

KMIN, KMAX = -149, 0
DTYPES = ("f32", "bf16", "f16")


# ---------------------------------------------------------------- backends
def make_backend(name):
    if name == "oracle":
        from oracle import po2_oracle as O

        def raw(vbits32, dtype, plus):
            v = vbits32.view(np.float32)
            R = lambda a: O._round_storage(a, dtype)
            with np.errstate(all="ignore"):
                if plus:
                    l = R(R(O.log2_cr_f32(R(v / np.float32(1.5)))) + np.float32(0.5))
                else:
                    l = R(O.log2_cr_f32(v))
                return np.rint(l).astype(np.float64)
        return raw

    import torch
    dev = "cuda" if name == "torch_cuda" else "cpu"
    if dev == "cuda":
        assert torch.cuda.is_available(), "torch_cuda backend needs a GPU"
    TD = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}

    def raw(vbits32, dtype, plus):
        v = torch.from_numpy(vbits32.view(np.int32).copy()).view(torch.float32).to(dev).to(TD[dtype])
        # same op sequence as utils/quantizers.py:26-27 / :46-47 (v is already abs(x/scale))
        if plus:
            r = torch.round(torch.log2(v / 1.5) + 0.5)
        else:
            r = torch.round(torch.log2(v))
        return r.double().cpu().numpy()
    return raw


# ---------------------------------------------------------------- grids
def grid_info(dtype):
    """(shift, max_index): index i on the dtype's positive grid <-> fp32 bits."""
    if dtype == "f32":
        return 0x3F800000
    if dtype == "bf16":
        return 0x3F80
    if dtype == "f16":
        return 0x3C00
    raise ValueError(dtype)


def idx_to_f32bits(idx, dtype):
    idx = np.asarray(idx, dtype=np.int64)
    if dtype == "f32":
        return idx.astype(np.uint32)
    if dtype == "bf16":
        return (idx.astype(np.uint32) << 16)
    h = idx.astype(np.uint16).view(np.float16).astype(np.float32)
    return h.view(np.uint32)


def scan(raw, dtype, plus, window):
    top = grid_info(dtype)            # index of 1.0
    ks = np.arange(KMIN, KMAX + 1)
    lo = np.zeros(ks.size, np.int64)          # raw(lo) < k   (index 0 is v = 0 -> -inf)
    hi = np.full(ks.size, top + 1, np.int64)  # sentinel: "never" if raw(1.0) < k
    # raw(top)
    rtop = raw(idx_to_f32bits(np.array([top]), dtype), dtype, plus)[0]
    hi_valid = ks <= rtop
    hi = np.where(hi_valid, top, top + 1)
    for _ in range(34):
        mid = (lo + hi) // 2
        active = (hi - lo) > 1
        if not active.any():
            break
        r = raw(idx_to_f32bits(np.minimum(mid, top), dtype), dtype, plus)
        ge = r >= ks
        hi = np.where(active & ge, mid, hi)
        lo = np.where(active & ~ge, mid, lo)
    # monotonicity check in a window around each boundary
    nonmono = []
    for k, b in zip(ks, hi):
        if b > top:
            continue
        a0 = max(1, b - window)
        a1 = min(top, b + window)
        idx = np.arange(a0, a1 + 1)
        r = raw(idx_to_f32bits(idx, dtype), dtype, plus)
        exp = idx >= b
        got = r >= k
        if not np.array_equal(exp, got):
            nonmono.append({"k": int(k), "first_bad_index": int(idx[np.flatnonzero(exp != got)[0]])})
    never = int(top + 1)
    bounds = [int(idx_to_f32bits(np.array([b]), dtype)[0]) if b <= top else 0xFFFFFFFF for b in hi]
    return {"kmin": KMIN, "kmax": KMAX, "bounds_f32bits": bounds, "non_monotone": nonmono,
            "never_index": never}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=("torch_cpu", "torch_cuda", "oracle"))
    ap.add_argument("--out")
    ap.add_argument("--window", type=int, default=4096)
    ap.add_argument("--diff", nargs=2)
    a = ap.parse_args()
    if a.diff:
        A, B = (json.load(open(p)) for p in a.diff)
        total = 0
        for key in sorted(A["tables"]):
            ta, tb = A["tables"][key]["bounds_f32bits"], B["tables"][key]["bounds_f32bits"]
            d = [(KMIN + i, x, y) for i, (x, y) in enumerate(zip(ta, tb)) if x != y]
            total += len(d)
            print(f"{key}: {len(d)} of {len(ta)} boundaries differ")
            for k, x, y in d:
                print(f"   k={k:5d}  {A['backend']}=0x{x:08X}  {B['backend']}=0x{y:08X}  delta={y - x:+d} ulp")
        print("total differing boundaries:", total)
        return
    raw = make_backend(a.backend)
    out = {"backend": a.backend, "tables": {}}
    if a.backend != "oracle":
        import torch
        out["torch"] = torch.__version__
        if a.backend == "torch_cuda":
            out["device"] = torch.cuda.get_device_name(0)
    for dtype in DTYPES:
        for plus in (False, True):
            key = f"{dtype}|{'po2+' if plus else 'po2'}"
            w = a.window if dtype == "f32" else 64
            t = scan(raw, dtype, plus, w)
            out["tables"][key] = t
            print(key, "non-monotone:", len(t["non_monotone"]),
                  "b[-1]=0x%08X b[0]=0x%08X" % (t["bounds_f32bits"][-2], t["bounds_f32bits"][-1]))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(out, open(a.out, "w"), indent=0)


if __name__ == "__main__":
    main()
