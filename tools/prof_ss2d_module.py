"""Minimal driver for ncu captures of one fused SS2D call (fwd+bwd) at the north-star shape: python tools/prof_ss2d_module.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import ceigm_unet_b200 as P
torch.manual_seed(0)
m = P.SS2D(d_model=96, d_state=16, ssm_ratio=2.0, k_group=4).cuda()
x = torch.randn(24, 56, 56, 96, device="cuda", requires_grad=True)
gy = torch.randn(24, 56, 56, 96, device="cuda")
for _ in range(2):
    y = m(x); y.backward(gy)
torch.cuda.synchronize()
print("ok")
