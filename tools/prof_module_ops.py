"""Which torch OPS launch the library kernels of one fused SS2D call (fwd+bwd): python tools/prof_module_ops.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import ceigm_unet_b200 as P
torch.set_float32_matmul_precision("medium")
torch.manual_seed(0)
m = P.SS2D(d_model=96, d_state=16, ssm_ratio=2.0, k_group=4).cuda()
x = torch.randn(24, 56, 56, 96, device="cuda", requires_grad=True)
gy = torch.randn(24, 56, 56, 96, device="cuda")
for _ in range(3):
    y = m(x); y.backward(gy)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    y = m(x); y.backward(gy)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.self_device_time_total) for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print("total device us", tot)
for k, c, t in rows[:45]:
    print("%-90s %4d %9.1f us %5.1f%%" % (k[:90], c, t, 100 * t / tot))
