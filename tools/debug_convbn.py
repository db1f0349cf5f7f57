"""Phase timestamps of the fused conv + norm kernel (PO2_TMA_DEBUG=16): python tools/debug_convbn.py [C HW K B]"""
import os, sys
os.environ["PO2_TMA_DEBUG"] = "16"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import po2_quantization_b200 as P
from po2_quantization_b200 import ops, fold
C, HW, K, B = (int(a) for a in (sys.argv[1:5] + ["16", "32", "16", "128"][len(sys.argv) - 1:]))
res = len(sys.argv) > 5
conv = P.QuantizedConv2d(C, K, 3, stride=1, padding=1, bias=False, quantize_fn=P.PowerOfTwoQuantizer, bits=4).cuda()
bn = P.FusedSyncBatchNorm(K).cuda().train()
x = torch.randn(B, C, HW, HW, device="cuda", requires_grad=True)
r = torch.randn(B, K, HW, HW, device="cuda") if res else None
ops.set_conv_mode("tf32")
P.conv_bn_act(conv, bn, x, r, True)
P.prefetch_weights([conv])
for _ in range(3):
    P.conv_bn_act(conv, bn, x, r, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    P.conv_bn_act(conv, bn, x, r, True)
e1.record()
torch.cuda.synchronize()
ws = list(fold._bn_workspaces.values())[0]
st = ws[200 * 1024:200 * 1024 + 48].view(torch.int64).cpu().tolist()
print("eager us per call", e0.elapsed_time(e1) * 1000 / 20)
names = ["pass0 done", "partials+fence", "barrier passed", "stats ready", "pass1 done"]
for i in range(1, 6):
    print(f"{names[i-1]:>16}: +{(st[i] - st[i-1]) / 1000:.2f} us  (t = {(st[i] - st[0]) / 1000:.2f})")
