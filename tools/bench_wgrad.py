"""Weight-gradient timings: po2::conv_wgrad_umma_kernel (+ reduce) vs aten.convolution_backward
(cuDNN, TF32 allowed = the reference's GPU default) on the ResNet-56 stride-1 3x3 shapes, batch 128.
CUDA-graph timed, 20 back-to-back repetitions (warm L2).

    python tools/bench_wgrad.py [--out gpurun_out/wgrad_layers.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po2_quantization_b200  # noqa: E402,F401
from po2_quantization_b200 import ops  # noqa: E402
from tools.bench_conv import graph_time  # noqa: E402

SHAPES = [("tiny 8->8 3x3 @32 B=2", 8, 32, 32, 8, 3, 1, 0), ("tiny 16->24 3x3 @8 B=6", 16, 8, 8, 24, 3, 1, 0),
          ("r56 16->16 3x3 @32", 16, 32, 32, 16, 3, 1, 18), ("r56 32->32 3x3 @16", 32, 16, 16, 32, 3, 1, 17),
          ("r56 64->64 3x3 @8", 64, 8, 8, 64, 3, 1, 17), ("mvit 128->64 3x3 @28 B=32", 128, 28, 28, 64, 3, 1, 0),
          ("mvit 32->128 1x1 @56 B=32", 32, 56, 56, 128, 1, 0, 0)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--profile", action="store_true", help="also print per-kernel durations (torch.profiler)")
    ap.add_argument("--compute", type=int, default=2, help="0: bf16-operand kernel, 2: tf32 (TMA-fed kernel where eligible)")
    a = ap.parse_args()
    from po2_quantization_b200 import _lib
    lib = _lib.load()
    REPS = 20
    rows = []
    for name, C, H, W, K, k, pad, cnt in SHAPES:
        B = 32 if "B=32" in name else (2 if "B=2" in name else (6 if "B=6" in name else 128))
        x = torch.randn(B, C, H, W, device="cuda")
        go = torch.randn(B, K, H, W, device="cuda")
        w = torch.randn(K, C, k, k, device="cuda")
        gw = torch.empty_like(w)

        def ours():
            assert ops.conv2d_wgrad_out(go, x, gw, pad, a.compute)

        kind = lib.po2_conv2d_wgrad_kernel_kind(B, C, H, W, K, k, k, 1, pad, 1, a.compute)
        ours()
        ref = torch.ops.aten.convolution_backward(go.double(), x.double(), w.double(), None, [1, 1], [pad, pad], [1, 1],
                                                  False, [0, 0], 1, [False, True, False])[1]
        err = ((gw.double() - ref).abs().max() / ref.abs().max()).item()

        def aten():
            torch.ops.aten.convolution_backward(go, x, w, None, [1, 1], [pad, pad], [1, 1], False, [0, 0], 1,
                                                [False, True, False])

        if a.profile:
            from torch.profiler import ProfilerActivity, profile
            ours()
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(10):
                    ours()
                torch.cuda.synchronize()
            for e in prof.key_averages():
                if e.device_time_total > 0:
                    print(f"   {name}: {e.key[:60]:60s} {e.device_time_total / e.count:8.2f} us x{e.count}")
        t_o = graph_time(ours, REPS) / REPS * 1e3
        t_a = graph_time(aten, REPS) / REPS * 1e3
        flops = 2.0 * B * H * W * C * K * k * k
        io = (x.numel() + go.numel()) * 4
        row = {"layer": name, "count_in_resnet56": cnt, "compute": a.compute, "kernel_kind": kind, "max_rel_err_vs_fp64": err,
               "po2_us": t_o, "aten_us": t_a, "speedup": t_a / t_o,
               "po2_TFLOPs": flops / t_o / 1e6, "po2_io_GBs": io / t_o / 1e3}
        rows.append(row)
        print(json.dumps(row))
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
