"""Run one quantized-conv shape a few times (for ncu captures).
    python tools/run_conv_once.py C H W K k stride pad groups [batch] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po2_quantization_b200  # noqa: E402,F401
from po2_quantization_b200 import ops  # noqa: E402

C, H, W, K, k, stride, pad, groups = (int(v) for v in sys.argv[1:9])
B = int(sys.argv[9]) if len(sys.argv) > 9 else 128
reps = int(sys.argv[10]) if len(sys.argv) > 10 else 3
x = torch.randn(B, C, H, W, device="cuda")
w = torch.randn(K, C // groups, k, k, device="cuda") * 0.1
y, codes, scale, _, _ = torch.ops.po2.quantize_full(w, 4, 1, True)
out = torch.empty(B, K, (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1, device="cuda")
for _ in range(reps):
    ops.conv2d_out(x, y, scale, out, stride, pad, groups, 0)
torch.cuda.synchronize()
print("ok", out.float().abs().mean().item())
