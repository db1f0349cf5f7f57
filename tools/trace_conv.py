"""Debug: per-role event timeline of the K3 conv kernel (first 4 CTAs), from a -DPO2_K3_TRACE build.
    python tools/trace_conv.py C H W K k stride pad [batch] [wgrad]"""
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
C, H, W, K, k, stride, pad = (int(v) for v in sys.argv[1:8])
B = int(sys.argv[8]) if len(sys.argv) > 8 else 128
lib_path = os.path.join(ROOT, "gpurun_out", "libpo2b200_trace.so")
os.makedirs(os.path.dirname(lib_path), exist_ok=True)
WGRAD = len(sys.argv) > 9 and sys.argv[9] == "wgrad"
COMPUTE = 2 if (len(sys.argv) > 9 and sys.argv[9] == "tf32") or (len(sys.argv) > 10 and sys.argv[10] == "tf32") else 0
src = [os.path.join(ROOT, "po2_quantization_b200", "csrc", f) for f in ("po2_quant.cu", "po2_conv.cu", "po2_conv_bwd.cu", "po2_bn.cu", "po2_lin.cu", "po2_sgd.cu")]
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared",
                       "-Xcompiler", "-fPIC", "-DPO2_K3_TRACE", *os.environ.get("PO2_TRACE_DEFS", "").split(), "-I", os.path.join(ROOT, "include"), "-o", lib_path, *src])
from po2_quantization_b200 import _lib  # noqa: E402
_lib.LIB_PATH = lib_path
import po2_quantization_b200  # noqa: E402,F401
from po2_quantization_b200 import ops  # noqa: E402
lib = _lib.load()
x = torch.randn(B, C, H, W, device="cuda")
w = torch.randn(K, C, k, k, device="cuda") * 0.1
y, codes, scale, _, _ = torch.ops.po2.quantize_full(w, 4, 1, True)
out = torch.empty(B, K, (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1, device="cuda")
gw = torch.empty_like(w)


def run():
    if WGRAD:
        assert ops.conv2d_wgrad_out(out, x, gw, pad, COMPUTE)
    else:
        ops.conv2d_out(x, y, scale, out, stride, pad, 1, COMPUTE)


out.normal_()
for _ in range(3):
    run()
torch.cuda.synchronize()
trace = torch.zeros(4 * 8 * 64, dtype=torch.int64, device="cuda")
lib.po2_debug_set_trace.argtypes = [ctypes.c_void_p]
assert lib.po2_debug_set_trace(trace.data_ptr()) == 0
run()
torch.cuda.synchronize()
t = trace.cpu().view(4, 8, 64)
names = ["mma0", "epi", "prod0", "prod1/bfull", "prod2", "prod3", "cta", "mma1"]
for cta in (0, 1):
    t0 = int(t[cta, 6, 0])
    print(f"--- CTA {cta} (SM clock cycles): setup done +{int(t[cta,6,1])-t0}, end +{int(t[cta,6,2])-t0}")
    for r in (0, 7, 1, 2, 3, 4, 5):
        ev = [(int(v) - t0) for v in t[cta, r] if v > 0]
        pairs = [f"{ev[i]}..{ev[i+1]}" for i in range(0, len(ev) - 1, 2)] or [str(e) for e in ev]
        if r in (4, 5):
            pairs = [str(e) for e in ev]
        print(f"  {names[r]:6s}", " ".join(pairs[:40] if r in (4, 5) else pairs[:12]))

