"""Kernel launch list of ONE fused K = 4 SS2D call (north-star shape), forward and backward separately, in launch order:
    python tools/launch_list_ss2d.py [fp32|tf32|bf16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import ceigm_unet_b200 as P
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
torch.manual_seed(0)
m = P.SS2D(d_model=96, d_state=16, ssm_ratio=2.0, k_group=4).cuda()
x = torch.randn(24, 56, 56, 96, device="cuda", requires_grad=True)
gy = torch.randn(24, 56, 56, 96, device="cuda")
ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16" else torch.autocast("cuda", enabled=False)
def fwd():
    with ctx:
        return m(x)
for _ in range(3):
    y = fwd(); y.backward(gy.to(y.dtype))
torch.cuda.synchronize()
for name, fn in (("forward", None), ("backward", None)):
    if name == "forward":
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            y = fwd(); torch.cuda.synchronize()
    else:
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            y.backward(gy.to(y.dtype)); torch.cuda.synchronize()
    evs = sorted([e for e in prof.events() if e.device_time_total > 0], key=lambda e: e.time_range.start)
    tot = sum(e.device_time_total for e in evs)
    print("## %s (%s): %d launches, %.1f us of GPU time" % (name, mode, len(evs), tot))
    for e in evs:
        print("  %8.1f us  %s" % (e.device_time_total, e.name[:130]))
