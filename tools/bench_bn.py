"""Per-kernel timings of the batch-norm path (csrc/po2_bn.cu) against torch's (cuDNN / native) batch
norm + add + ReLU on the ResNet-56 activation shapes, batch 128.

    python tools/bench_bn.py [--out gpurun_out/bn_layers.json]

Every candidate is captured in a CUDA graph; "warm" = 20 back-to-back repetitions (the 8.4 MB
tensors stay L2-resident, as they do behind a conv in the real step).  GB/s = algorithmic bytes
(forward: read x, write y [+ read residual]; backward: read dy, x, y, write dx [+ write dres]) per
the summed time of the two kernels of that direction, against MEASURED_PEAKS.json's HBM copy peak
for orientation only -- these passes run largely out of L2.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po2_quantization_b200 as P  # noqa: E402
from tools.bench_conv import graph_time  # noqa: E402

SHAPES = [("r56 stage1 16ch @32x32", 128, 16, 32, 32, 19), ("r56 stage2 32ch @16x16", 128, 32, 16, 16, 19),
          ("r56 stage3 64ch @8x8", 128, 64, 8, 8, 19), ("mvit 64ch @128x128 B=32", 32, 64, 128, 128, 0)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    REPS = 20
    rows = []
    for name, B, C, H, W, cnt in SHAPES:
        x = torch.randn(B, C, H, W, device="cuda")
        res = torch.randn(B, C, H, W, device="cuda")
        go = torch.randn(B, C, H, W, device="cuda")
        mine = P.FusedSyncBatchNorm(C).cuda().train()
        stock = torch.nn.BatchNorm2d(C).cuda().train()
        nbytes = x.numel() * 4
        row = {"shape": name, "tensor_MB": nbytes / 1e6, "count_in_resnet56": cnt}
        for label, with_res in (("bn_relu", False), ("bn_add_relu", True)):
            def f_mine(xi, ri):
                return mine(xi, ri, True)

            def f_stock(xi, ri):
                y = stock(xi)
                if ri is not None:
                    y = y + ri
                return F.relu(y)

            for who, f in (("po2", f_mine), ("torch", f_stock)):
                def fwd():
                    with torch.no_grad():
                        f(x, res if with_res else None)                         # noqa: B023

                def fb():
                    # fresh leaves on the capturing stream (a leaf made on the default stream would make
                    # autograd synchronise with that stream, which a capture forbids)
                    xi = x.detach().requires_grad_(True)
                    ri = res.detach().requires_grad_(True) if with_res else None    # noqa: B023
                    y2 = f(xi, ri)                                              # noqa: B023
                    torch.autograd.grad(y2, (xi, ri) if ri is not None else (xi,), go)

                t_f = graph_time(fwd, REPS) / REPS * 1e3
                t_fb = graph_time(fb, REPS) / REPS * 1e3
                row[f"{label}_{who}_fwd_us"] = t_f
                row[f"{label}_{who}_fwd_bwd_us"] = t_fb
            fw_bytes = nbytes * (3 if with_res else 2)
            bw_bytes = nbytes * (5 if with_res else 4)
            row[f"{label}_po2_fwd_GBs"] = fw_bytes / (row[f"{label}_po2_fwd_us"] * 1e-6) / 1e9
            row[f"{label}_po2_bwd_GBs"] = bw_bytes / ((row[f"{label}_po2_fwd_bwd_us"] - row[f"{label}_po2_fwd_us"]) * 1e-6) / 1e9
        rows.append(row)
        print(json.dumps(row))
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
