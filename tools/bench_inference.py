"""PTQ inference forward of the BASELINE.json configs that are parity cases rather than the bench line:
ResNet-20 PO2+ 4-bit (configs[0]), MobileNetV2 PO2+ 4-bit (configs[2]), MobileViT-xs PO2+ 8-bit
(configs[3]; 224x224 patch (1,1) and 256x256 patch (2,2), batch 32 here) -- images/s of the whole
model forward (eval mode, CUDA graph), po2 tensor-core path vs the same model with the convs on cuDNN.

    python tools/bench_inference.py [--out gpurun_out/inference.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po2_quantization_b200 as P  # noqa: E402
from po2_quantization_b200 import ops  # noqa: E402
from workloads import mobilenet_v2_cifar, mobilevit_xs, resnet_cifar  # noqa: E402


def graph_ms(fn, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    cases = [
        ("resnet20 po2+ 4b PTQ, B=128, 32x32", lambda: resnet_cifar(20, 10, None, 4), 4, 128, (32, 32)),
        ("resnet56 po2+ 4b PTQ, B=128, 32x32", lambda: resnet_cifar(56, 10, None, 4), 4, 128, (32, 32)),
        ("mobilenetv2 po2+ 4b PTQ, B=128, 32x32", lambda: mobilenet_v2_cifar(10, None, 4), 4, 128, (32, 32)),
        ("mobilevit-xs po2+ 8b PTQ, B=32, 224x224 patch(1,1)", lambda: mobilevit_xs((224, 224), 1000, (1, 1), None, 8), 8, 32, (224, 224)),
        ("mobilevit-xs po2+ 8b PTQ, B=32, 256x256 patch(2,2)", lambda: mobilevit_xs((256, 256), 1000, (2, 2), None, 8), 8, 32, (256, 256)),
    ]
    rows = []
    for name, build, bits, B, img in cases:
        torch.manual_seed(8)
        m = build().cuda().eval()
        P.quantize_model(m, P.PowerOfTwoPlusQuantizer, bits)
        x = torch.randn(B, 3, *img, device="cuda")
        r = {"case": name, "batch": B}
        for mode in ("tc", "cudnn"):
            ops.set_conv_mode(mode)
            ms = graph_ms(lambda: m(x))
            r[f"ms_{mode}"] = ms
            r[f"images_per_s_{mode}"] = B / ms * 1e3
        ops.set_conv_mode(ops.DEFAULT_CONV_MODE)
        r["speedup_vs_cudnn_convs"] = r["ms_cudnn"] / r["ms_tc"]
        rows.append(r)
        print(json.dumps(r), flush=True)
        del m, x
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
