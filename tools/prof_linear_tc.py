"""Minimal driver for ncu captures of the tensor-core projection kernel at the north-star SS2D shape
(d_model 96 -> 2 x 192 and 192 -> 96 on a batch-24 56 x 56 map): python tools/prof_linear_tc.py [fp32|bf16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ceigm_unet_b200 import ops
dtype = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
Bn, H, W, C, D = 24, 56, 56, 96, 192
x = torch.randn(Bn, H, W, C, device="cuda").to(dtype)
Win = (torch.randn(2 * D, C, device="cuda") / C ** 0.5).to(dtype)
Wout = (torch.randn(C, D, device="cuda") / D ** 0.5).to(dtype)
y = torch.randn(Bn, H, W, D, device="cuda").to(dtype)
for _ in range(2):
    ops.linear_tc(x, Win, None, [(D, ("planes", H * W), False), (D, "rows", False)])
    ops.linear_tc(y, Wout, None, [(C, "rows", False)])
torch.cuda.synchronize()
print("ok")
