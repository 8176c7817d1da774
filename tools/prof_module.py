"""Kernel-level time breakdown (torch profiler) of one fused SS2D call, fwd+bwd, in the VMamba / north-star regime.
    python tools/prof_module.py [d_model] [d_state] [batch] [H]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import ceigm_unet_b200 as P
a = [int(v) for v in sys.argv[1:]]
d_model, N, Bn, H = (a + [96, 16, 24, 56][len(a):])[:4]
torch.set_float32_matmul_precision("medium")      # what the reference trains with (train_synapse.py:21)
torch.manual_seed(0)
m = P.SS2D(d_model=d_model, d_state=N, ssm_ratio=2.0, k_group=4).cuda()
x = torch.randn(Bn, H, H, d_model, device="cuda", requires_grad=True)
gy = torch.randn(Bn, H, H, d_model, device="cuda")
for _ in range(3):
    y = m(x); y.backward(gy)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        y = m(x); y.backward(gy)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
