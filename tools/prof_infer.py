#!/usr/bin/env python
"""Kernel-time breakdown of one GM-UNet inference step (512^2, batch 64, bf16, fused modules + harness rebindings) by
torch.profiler. python tools/prof_infer.py > profiles/..."""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from harness import graph_step, workloads as W  # noqa: E402

dev = torch.device("cuda", 0)
net = W.build(9, "fused", dev).eval()
graph_step.make_capturable()
x, _ = W.synthetic_batch(64, 512, 9, seed=7)


@torch.no_grad()
def one():
    xd = x.to(dev, non_blocking=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = net(xd)
    return torch.argmax(torch.softmax(logits.float(), dim=1), dim=1).to(torch.uint8)


for _ in range(2):
    one()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    one()
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        tot[ev.name[:100]][0] += 1
        tot[ev.name[:100]][1] += ev.device_time
total = sum(v[1] for v in tot.values())
print("# fused inference 512^2 batch 64: %d GPU kernels/memcpys, %.2f ms of GPU time" % (sum(v[0] for v in tot.values()), total / 1e3))
print("%-100s %6s %9s %6s" % ("kernel", "n", "ms", "%"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%-100s %6d %9.3f %6.2f" % (k, v[0], v[1] / 1e3, 100 * v[1] / total))
