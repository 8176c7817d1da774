"""Minimal driver for ncu captures of the fused epilogue + out_proj kernel at the north-star shape: python tools/prof_gate_proj.py [bf16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ceigm_unet_b200 import ops
dtype = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
Bn, K, D, H, W, C = 24, 4, 192, 56, 56, 96
L = H * W
ys = torch.randn(Bn, K, D, L, device="cuda")
lnw, lnb = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
z = torch.randn(Bn, L, D, device="cuda").to(dtype)
Wt = (torch.randn(C, D, device="cuda") / D ** 0.5).to(dtype)
for want_g in (False, True):
    for _ in range(2):
        ops.gate_proj_fwd(ys, lnw, lnb, z, True, 1e-5, Wt, None, (H, W), 0b1010, want_g)
torch.cuda.synchronize()
print("ok")
