import sys
sys.path.insert(0, "/root/repo")
import torch
from ceigm_unet_b200 import ops
Bn, K, D, H, W, C = 24, 1, 192, 56, 56, 96
L = H * W
ys = torch.randn(Bn, K, D, L, device="cuda")
lnw, lnb = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
Wt = torch.randn(C, D, device="cuda") / D ** 0.5
for _ in range(3):
    ops.gate_proj_fwd(ys, lnw, lnb, None, True, 1e-5, Wt, None, (H, W), 0, False)
torch.cuda.synchronize()
print("ok")
