"""Run the backward / norm kernels of one ResNet-56 layer shape a few times (for ncu captures).
    python tools/run_step_kernels_once.py [C H W K] [batch]
Launches: conv_wgrad_tma_kernel + conv_wgrad_reduce_kernel, bn_fwd_fused_kernel, bn_bwd_fused_kernel (one rank)
on a (batch, C, H, W) activation."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po2_quantization_b200 as P  # noqa: E402
from po2_quantization_b200 import ops  # noqa: E402

C, H, W, K = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (16, 32, 32, 16)
B = int(sys.argv[5]) if len(sys.argv) > 5 else 128
x = torch.randn(B, C, H, W, device="cuda")
go = torch.randn(B, K, H, W, device="cuda")
gw = torch.empty(K, C, 3, 3, device="cuda")
bn = P.FusedSyncBatchNorm(C).cuda().train()
res = torch.randn_like(x)
for _ in range(3):
    assert ops.conv2d_wgrad_out(go, x, gw, 1, 2)          # tf32 mode: conv_wgrad_tma_kernel (K5T)
    xi = x.detach().requires_grad_(True)
    y = bn(xi, res, True)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok", gw.abs().mean().item(), y.mean().item())
