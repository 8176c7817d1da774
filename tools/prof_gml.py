"""Kernel-level time breakdown (torch profiler, CUDA activities) of one GroupMambaLayer fwd+bwd at a stage shape.
    python tools/prof_gml.py [C] [H] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import ceigm_unet_b200 as P
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 56
Bn = int(sys.argv[3]) if len(sys.argv) > 3 else 24
torch.set_float32_matmul_precision("medium")      # what the reference trains with (train_synapse.py:21)
torch.manual_seed(0)
m = P.GroupMambaLayer(C, C).cuda()
x = torch.randn(Bn, H * H, C, device="cuda", requires_grad=True)
gy = torch.randn(Bn, H * H, C, device="cuda")
for _ in range(3):
    y = m(x, H, H); y.backward(gy)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        y = m(x, H, H); y.backward(gy)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
