"""One eager PTQ forward of a workload model between cudaProfilerStart/Stop (for an ncu launch list).
    python tools/profile_inference.py mobilenet|resnet20|mobilevit [tc|cudnn]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po2_quantization_b200 as P  # noqa: E402
from po2_quantization_b200 import ops  # noqa: E402
from workloads import mobilenet_v2_cifar, mobilevit_xs, resnet_cifar  # noqa: E402

name = sys.argv[1]
ops.set_conv_mode(sys.argv[2] if len(sys.argv) > 2 else "tc")
torch.manual_seed(8)
if name == "mobilenet":
    m, bits, x = mobilenet_v2_cifar(10, None, 4), 4, torch.randn(128, 3, 32, 32)
elif name == "resnet20":
    m, bits, x = resnet_cifar(20, 10, None, 4), 4, torch.randn(128, 3, 32, 32)
else:
    m, bits, x = mobilevit_xs((224, 224), 1000, (1, 1), None, 8), 8, torch.randn(32, 3, 224, 224)
m = m.cuda().eval()
x = x.cuda()
P.quantize_model(m, P.PowerOfTwoPlusQuantizer, bits)
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok")
