"""Compile quantize_kernel variants (-DPO2_Q_UNROLL / _MINBLOCKS / _CTAS_PER_SM) and time pass 2 at 2^28.
    python tools/tune_quantizer.py"""
import ctypes
import itertools
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = [os.path.join(ROOT, "po2_quantization_b200", "csrc", f) for f in ("po2_quant.cu", "po2_conv.cu")]
out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
variants = [(4, 1, 8), (4, 6, 8), (4, 8, 8), (8, 1, 8), (8, 4, 8), (2, 8, 8), (4, 4, 16), (8, 6, 6), (4, 5, 5), (6, 5, 5)]
n = 1 << 28
res = []
for unroll, minb, cps in variants:
    so = os.path.join(out_dir, f"libq_u{unroll}_m{minb}_c{cps}.so")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           f"-DPO2_Q_UNROLL={unroll}", f"-DPO2_Q_MINBLOCKS={minb}", f"-DPO2_Q_CTAS_PER_SM={cps}",
           "-I", os.path.join(ROOT, "include"), "-o", so, *src]
    subprocess.check_call(cmd)
    lib = ctypes.CDLL(so)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.po2_quantize.restype = i32
    lib.po2_quantize.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp]
    row = {"unroll": unroll, "minblocks": minb, "ctas_per_sm": cps}
    for dt, code, es in ((torch.float32, 0, 4), (torch.bfloat16, 1, 2)):
        x = torch.randn(n, device="cuda").to(dt)
        y = torch.empty_like(x)
        s = x.abs().max().float().reshape(())
        st = torch.cuda.current_stream().cuda_stream
        fn = lambda: lib.po2_quantize(x.data_ptr(), y.data_ptr(), None, None, None, s.data_ptr(), n, code, 4, 1, 0, 0, st)
        for _ in range(3):
            assert fn() == 0
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        row[str(dt).split(".")[-1]] = round(2 * es * n / ms / 1e6, 1)
        del x, y
    res.append(row)
    print(json.dumps(row), flush=True)
json.dump(res, open(os.path.join(out_dir, "tune_quantizer.json"), "w"), indent=1)
