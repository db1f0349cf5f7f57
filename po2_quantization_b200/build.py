"""Build libpo2b200.so in-tree with nvcc for sm_100a (B200).  No torch headers are involved:
the library is plain CUDA behind the C ABI in include/po2_b200.h.

    python -m po2_quantization_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpo2b200.so")
SOURCES = ["po2_quant.cu", "po2_conv.cu", "po2_conv_bwd.cu", "po2_bn.cu", "po2_lin.cu", "po2_sgd.cu"]
HEADERS = ["po2_common.cuh", "po2_conv_tma.cuh", "po2_wgrad_tma.cuh", "po2_boundaries.inc", os.path.join(ROOT, "include", "po2_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libpo2b200.so cannot be built")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into one shared library; returns its path."""
    if not force and not _stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           "-I", os.path.join(ROOT, "include"), "-o", LIB, *srcs, "-lcuda"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
