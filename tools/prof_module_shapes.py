import os, sys
sys.path.insert(0, "/root/repo")
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    y = m(x); y.backward(gy)
    torch.cuda.synchronize()
for e in prof.events():
    if e.name in ("aten::copy_", "aten::add_", "aten::sum", "aten::mm", "aten::bmm") and e.device_time_total > 8:
        st = [s for s in (e.stack or []) if "ceigm" in s or "repo" in s][:2]
        print("%-12s %8.1f us %s %s" % (e.name, e.device_time_total, e.input_shapes, st))
