"""One tf32 conv through the TMA-fed kernel (for compute-sanitizer / ncu): python tools/run_tma_once.py C H W K k [B]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po2_quantization_b200  # noqa: E402,F401
from po2_quantization_b200 import ops, _lib  # noqa: E402

C, H, W, K, k = [int(v) for v in sys.argv[1:6]]
B = int(sys.argv[6]) if len(sys.argv) > 6 else 128
pad = 1 if k == 3 else 0
x = torch.randn(B, C, H, W, device="cuda")
w = torch.randn(K, C, k, k, device="cuda") * 0.1
y, codes, scale, _, _ = torch.ops.po2.quantize_full(w, 4, 1, True)
out = torch.empty(B, K, H, W, device="cuda")
print("kind", _lib.load().po2_conv2d_kernel_kind(B, C, H, W, K, k, k, 1, pad, 1, 2), flush=True)
ops.conv2d_out(x, y, scale, out, 1, pad, 1, 2)
torch.cuda.synchronize()
torch.backends.cudnn.allow_tf32 = False
ref = F.conv2d(x, y, None, 1, pad)
err = (out - ref).abs()
print("max rel err", (err.max() / ref.abs().max()).item(), "bad frac", (err > 1e-2 * ref.abs().max()).float().mean().item())
