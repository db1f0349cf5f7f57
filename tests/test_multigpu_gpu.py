"""Cross-GPU checks as part of the `-m gpu` suite; they skip themselves on a box with a single GPU.
One process per GPU under torchrun (tools/check_sync_bn.py): FusedSyncBatchNorm's in-kernel NVLink
exchange and its NCCL path against whole-batch batch norm in fp64 (eager and under CUDA-graph replay),
BatchSharded's bucketed gradient averaging against plain all_reduce, and sharded_quantize on the
sm_100a kernels against the single-device quantizer (bit-exact)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 8])
def test_cross_gpu_sync_bn_gradient_averaging_and_sharded_quantizer(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_sync_bn.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    modes = {l["requested"]: l for l in lines if "requested" in l}
    assert modes["peer"]["max_rel_err"] < 5e-5 and modes["nccl"]["max_rel_err"] < 5e-5
    assert modes["peer"]["timeout_flag"] == 0
    assert any(l.get("sharded_quantize_cuda_backend_mismatching_cases") == 0 for l in lines)
    out = os.path.join(os.environ.get("GRAFT_REPO_ROOT", ROOT), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        open(os.path.join(out, f"check_cross_gpu_n{world}.log"), "w").write(r.stdout)
    except OSError:
        pass
