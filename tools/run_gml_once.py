"""One GroupMambaLayer fwd+bwd at a stage shape (driver for ncu captures).  python tools/run_gml_once.py [C] [H] [batch] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import ceigm_unet_b200 as P
a = [int(v) for v in sys.argv[1:]]
C, H, Bn, iters = (a + [64, 56, 24, 2][len(a):])[:4]
torch.set_float32_matmul_precision("medium")
torch.manual_seed(0)
m = P.GroupMambaLayer(C, C).cuda()
x = torch.randn(Bn, H * H, C, device="cuda", requires_grad=True)
gy = torch.randn(Bn, H * H, C, device="cuda")
for _ in range(iters):
    y = m(x, H, H); y.backward(gy)
torch.cuda.synchronize()
print("ok")
