"""Standalone PO2/PO2+ quantizer sweep (BASELINE.json configs[4]): GB/s vs measured HBM peak.

    python tools/bench_quantizer.py [--max-log2 30] [--iters 20] [--out gpurun_out/quant_sweep.json]

Algorithmic bytes per element: absmax read + quantize read + dequantized write
= 3 * elem_size (+ bits/8 with codes).  CUDA events on the current stream; L2 is flushed between
iterations when the tensor is smaller than 256 MB.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from po2_quantization_b200 import ops  # noqa: E402


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def time_fn(fn, iters, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for a, b in ev:
        if flush is not None:
            flush.add_(1)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=20)
    ap.add_argument("--max-log2", type=int, default=30)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--bits", type=int, nargs="*", default=[4])
    a = ap.parse_args()
    peak, src = peak_gbs()
    flush_buf = torch.zeros(320 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")
    rows = []
    for dt, es in ((torch.float32, 4), (torch.bfloat16, 2)):
        for lg in range(a.min_log2, a.max_log2 + 1, 2):
            n = 1 << lg
            x = torch.randn(n, device="cuda", dtype=torch.float32).to(dt) if n <= (1 << 30) else \
                torch.cat([torch.randn(1 << 30, device="cuda").to(dt) for _ in range(n >> 30)])
            y = torch.empty_like(x)
            s = torch.empty((), dtype=torch.float32, device="cuda")
            codes = torch.empty((n + 1) // 2, dtype=torch.uint8, device="cuda")
            flush = flush_buf if n * es < 256 * 1024 * 1024 else None
            for plus in (False, True):
                for bits in a.bits:
                    med, best = time_fn(lambda: ops.quantize_fused_out(x, y, s, bits, 1, plus), a.iters, flush)
                    med_c, _ = time_fn(lambda: ops.quantize_fused_out(x, y, s, bits, 1, plus, codes=codes), a.iters, flush)
                    t_abs, _ = time_fn(lambda: ops.absmax_out(x, s), a.iters, flush)
                    t_q, _ = time_fn(lambda: ops.quantize_out(x, y, s, bits, 1, plus), a.iters, flush)
                    row = {"dtype": str(dt).split(".")[-1], "log2n": lg, "quantizer": "po2+" if plus else "po2",
                           "bits": bits, "ms_fused": med, "ms_fused_best": best,
                           "GBs_alg": 3 * es * n / med / 1e6, "frac_of_%s_peak" % src: 3 * es * n / med / 1e6 / peak,
                           "ms_with_codes": med_c, "GBs_alg_with_codes": (3 * es + bits / 8) * n / med_c / 1e6,
                           "ms_absmax": t_abs, "GBs_absmax": es * n / t_abs / 1e6,
                           "ms_quantize": t_q, "GBs_quantize": 2 * es * n / t_q / 1e6}
                    rows.append(row)
                    print(json.dumps(row), flush=True)
            del x, y, codes
    if a.out:
        json.dump({"peak_gbs": peak, "peak_source": src, "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
