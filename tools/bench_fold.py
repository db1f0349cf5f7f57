"""conv2d_packed vs conv2d_packed_ep (folded BatchNorm + residual + ReLU) per ResNet layer class, CUDA-graph timed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po2_quantization_b200  # noqa: E402,F401
from po2_quantization_b200 import ops  # noqa: E402
from tools.bench_conv import graph_time  # noqa: E402

for C, HW, K in ((16, 32, 16), (32, 16, 32), (64, 8, 64)):
    x = torch.randn(128, C, HW, HW, device="cuda")
    w = torch.randn(K, C, 3, 3, device="cuda") * 0.1
    y, _, scale, _, _ = torch.ops.po2.quantize_full(w, 4, 1, True)
    packed = ops.conv2d_pack(y, scale, x.shape, 1, 1, 1, 2)
    a, b = torch.rand(K, device="cuda") + 0.5, torch.randn(K, device="cuda")
    res = torch.randn(128, K, HW, HW, device="cuda")
    t0 = graph_time(lambda: torch.ops.po2.conv2d_packed(x, packed, scale, K, 3, 3, 1, 1, 1, 2), 20) / 20 * 1e3
    t1 = graph_time(lambda: torch.ops.po2.conv2d_packed_ep(x, packed, scale, K, 3, 3, 1, 1, 1, 2, a, b, None, 1), 20) / 20 * 1e3
    t2 = graph_time(lambda: torch.ops.po2.conv2d_packed_ep(x, packed, scale, K, 3, 3, 1, 1, 1, 2, a, b, res, 1), 20) / 20 * 1e3
    print(f"{C}->{K} @{HW}: plain {t0:.2f} us, +bn+relu {t1:.2f} us, +bn+res+relu {t2:.2f} us")
