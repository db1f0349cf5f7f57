"""Debug: phase stamps (globaltimer, ns) of the one-launch BatchNorm kernels, CTA (0, 0), from a -DPO2_BN_TRACE build.
    python tools/trace_bn.py [B C H] [res]"""
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
args = [a for a in sys.argv[1:] if a != "res"]
B, C, H = (int(v) for v in (args + ["128", "16", "32"][len(args):])[:3])
RES = "res" in sys.argv
lib_path = os.path.join(ROOT, "gpurun_out", "libpo2b200_bntrace.so")
os.makedirs(os.path.dirname(lib_path), exist_ok=True)
src = [os.path.join(ROOT, "po2_quantization_b200", "csrc", f) for f in
       ("po2_quant.cu", "po2_conv.cu", "po2_conv_bwd.cu", "po2_bn.cu", "po2_lin.cu", "po2_sgd.cu")]
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                       "-DPO2_BN_TRACE", "-I", os.path.join(ROOT, "include"), "-o", lib_path, *src, "-lcuda"])
from po2_quantization_b200 import _lib  # noqa: E402
_lib.LIB_PATH = lib_path
import po2_quantization_b200 as P  # noqa: E402
lib = _lib.load()
bn = P.FusedSyncBatchNorm(C).cuda().train()
x = torch.randn(B, C, H, H, device="cuda", requires_grad=True)
r = torch.randn(B, C, H, H, device="cuda", requires_grad=True) if RES else None
go = torch.randn(B, C, H, H, device="cuda")
for _ in range(3):
    bn(x, r, True).backward(go)
torch.cuda.synchronize()
trace = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.po2_debug_set_bn_trace.argtypes = [ctypes.c_void_p]
assert lib.po2_debug_set_bn_trace(trace.data_ptr()) == 0
bn(x, r, True).backward(go)
torch.cuda.synchronize()
t = trace.cpu().tolist()
names = ["loads + local sums", "block sum + partial", "channel barrier", "combine statistics", "apply + stores issued"]
for base, what in ((0, "forward"), (8, "backward")):
    print(f"--- {what} B={B} C={C} H={H} residual={RES} (CTA (0,0), microseconds)")
    for i in range(5):
        print(f"  {names[i]:>24}: +{(t[base + i + 1] - t[base + i]) / 1000:.2f}   (t = {(t[base + i + 1] - t[base]) / 1000:.2f})")
