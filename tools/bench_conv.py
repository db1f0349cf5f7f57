"""Per-layer quantized-conv forward timings: po2 tensor-core path vs cuDNN (TF32 = the reference's
default on a GPU; fp32) for the layer shapes of SURVEY.md section 8a.

    python tools/bench_conv.py [--batch 128] [--out gpurun_out/conv_layers.json]

Each candidate is captured in a CUDA graph (so Python/launch overhead is not what is measured);
"cold" = graph of [rewrite a 320 MB buffer; conv] minus graph of [rewrite the buffer], i.e. the conv
starts with its operands out of L2; "warm" = 20 back-to-back convs, operands L2-resident where they
fit.  po2 timings include the weight-pack kernel that runs before every conv in QAT mode.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po2_quantization_b200  # noqa: E402,F401
from po2_quantization_b200 import ops  # noqa: E402

SHAPES = [  # name, C, H, W, K, k, stride, pad, groups, count in ResNet-56 (0 = other model)
    ("r56 16->16 3x3 s1 @32", 16, 32, 32, 16, 3, 1, 1, 1, 18),
    ("r56 16->32 3x3 s2 @32", 16, 32, 32, 32, 3, 2, 1, 1, 1),
    ("r56 16->32 1x1 s2 @32", 16, 32, 32, 32, 1, 2, 0, 1, 1),
    ("r56 32->32 3x3 s1 @16", 32, 16, 16, 32, 3, 1, 1, 1, 17),
    ("r56 32->64 3x3 s2 @16", 32, 16, 16, 64, 3, 2, 1, 1, 1),
    ("r56 32->64 1x1 s2 @16", 32, 16, 16, 64, 1, 2, 0, 1, 1),
    ("r56 64->64 3x3 s1 @8", 64, 8, 8, 64, 3, 1, 1, 1, 17),
    ("mbv2 pw 16->96 @16", 16, 16, 16, 96, 1, 1, 0, 1, 0),
    ("mbv2 pw 144->32 @4", 144, 4, 4, 32, 1, 1, 0, 1, 0),
    ("mbv2 pw 960->320 @1", 960, 1, 1, 320, 1, 1, 0, 1, 0),
    ("mbv2 dw 96 s2 @16", 96, 16, 16, 96, 3, 2, 1, 96, 0),
    ("mbv2 dw 384 s1 @2", 384, 2, 2, 384, 3, 1, 1, 384, 0),
    ("mvit 128->64 3x3 @28", 128, 28, 28, 64, 3, 1, 1, 1, 0),
    ("mvit 32->128 1x1 @112", 32, 112, 112, 128, 1, 1, 0, 1, 0),
]


MOBILENET_PW = [  # the 33 pointwise layers of MobileNetV2-CIFAR (C, HW side, K, count)
    (32, 16, 16, 1), (16, 16, 96, 1), (96, 8, 24, 1), (24, 8, 144, 2), (144, 8, 24, 1), (144, 4, 32, 1), (32, 4, 192, 3),
    (192, 4, 32, 2), (192, 2, 64, 1), (64, 2, 384, 4), (384, 2, 64, 3), (384, 2, 96, 1), (96, 2, 576, 3), (576, 2, 96, 2),
    (576, 1, 160, 1), (160, 1, 960, 3), (960, 1, 160, 2), (960, 1, 320, 1)]
MOBILENET_DW = [(32, 16, 1, 1), (96, 16, 2, 1), (144, 8, 1, 1), (144, 8, 2, 1), (192, 4, 1, 2), (192, 4, 2, 1), (384, 2, 1, 4),
                (576, 2, 1, 2), (576, 2, 2, 1), (960, 1, 1, 3)]


def graph_time(body, reps, iters=5):
    """median ms of one replay of a graph that runs `body` `reps` times"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body()
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--out", default=None)
    ap.add_argument("--compute", type=int, default=0, help="0: bf16 operands, 2: tf32 operands (TMA-fed where eligible)")
    ap.add_argument("--only", default=None, help="substring filter on the layer name")
    ap.add_argument("--set", default="default", help="default | mobilenet (all 33 pointwise + 17 depthwise layers)")
    a = ap.parse_args()
    from po2_quantization_b200 import _lib
    lib = _lib.load()
    flush = torch.zeros(320 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")
    REPS = 10
    t_flush = graph_time(lambda: flush.add_(1), REPS)
    rows = []
    shapes = SHAPES
    if a.set == "mobilenet":
        shapes = [(f"mbv2 pw {C}->{K} @{hw}", C, hw, hw, K, 1, 1, 0, 1, n) for C, hw, K, n in MOBILENET_PW] + \
                 [(f"mbv2 dw {C} s{st} @{hw}", C, hw, hw, C, 3, st, 1, C, n) for C, hw, st, n in MOBILENET_DW]
    for name, C, H, W, K, k, stride, pad, groups, cnt in shapes:
        if a.only and a.only not in name:
            continue
        B = a.batch
        x = torch.randn(B, C, H, W, device="cuda")
        w = torch.randn(K, C // groups, k, k, device="cuda") * 0.1
        y, codes, scale, _, _ = torch.ops.po2.quantize_full(w, 4, 1, True)
        P, Q = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        out = torch.empty(B, K, P, Q, device="cuda")
        flops = 2.0 * B * K * P * Q * (C // groups) * k * k
        bytes_io = 4.0 * (x.numel() + out.numel())
        r = {"layer": name, "batch": B, "gflop": flops / 1e9, "io_mb": bytes_io / 1e6, "count_r56": cnt,
             "compute": a.compute,
             "kernel_kind": lib.po2_conv2d_kernel_kind(B, C, H, W, K, k, k, stride, pad, groups, a.compute)}
        cands = {"po2_tc": lambda: ops.conv2d_out(x, y, scale, out, stride, pad, groups, a.compute)}
        packed = ops.conv2d_pack(y, scale, x.shape, stride, pad, groups, a.compute)
        if packed is not None:     # static weights (PTQ) / multi-tensor prefetch (QAT): the conv kernel alone
            cands["po2_packed"] = lambda: torch.ops.po2.conv2d_packed(x, packed, scale, K, k, k, stride, pad, groups, a.compute)
        torch.backends.cudnn.allow_tf32 = False
        ref = F.conv2d(x, y, None, stride, pad, 1, groups)
        cands["po2_tc"]()
        r["max_rel_err_vs_fp32"] = ((out - ref).abs().max() / ref.abs().max()).item()
        torch.backends.cudnn.allow_tf32 = True
        F.conv2d(x, y, None, stride, pad, 1, groups)
        cands["cudnn_tf32"] = lambda: F.conv2d(x, y, None, stride, pad, 1, groups)
        for key, fn in cands.items():
            def cold():
                flush.add_(1)
                fn()
            r[f"us_{key}_cold"] = (graph_time(cold, REPS) - t_flush) / REPS * 1e3
            r[f"us_{key}_warm"] = graph_time(fn, 20) / 20 * 1e3
        r["tflops_po2_tc_cold"] = flops / r["us_po2_tc_cold"] / 1e6
        r["io_GBs_po2_tc_cold"] = bytes_io / r["us_po2_tc_cold"] / 1e3
        r["io_GBs_cudnn_cold"] = bytes_io / r["us_cudnn_tf32_cold"] / 1e3
        rows.append(r)
        r["max_rel_err_vs_fp32"] = ((out.double() - ref.double()).abs().max() / ref.double().abs().max()).item()
        print(json.dumps({k_: (round(v, 5 if "err" in k_ else 2) if isinstance(v, float) else v) for k_, v in r.items()}), flush=True)
    tot = lambda key: sum(r[key] * r["count_r56"] for r in rows)
    tot = lambda key: sum(r.get(key, 0.0) * r["count_r56"] for r in rows)
    summ = {"resnet56_forward_qconv_us": {k_: round(tot(k_), 1) for k_ in
            ("us_po2_tc_cold", "us_po2_tc_warm", "us_po2_packed_cold", "us_po2_packed_warm", "us_cudnn_tf32_cold",
             "us_cudnn_tf32_warm")}}
    print(json.dumps(summ))
    if a.out:
        json.dump({"rows": rows, **summ}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
