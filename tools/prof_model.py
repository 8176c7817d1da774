#!/usr/bin/env python
"""Kernel-time breakdown of one GM-UNet training step (224^2, batch 24, bf16) by torch.profiler: which kernels the GPU time of
the graphed step is made of. python tools/prof_model.py [dropin|fused] [capturable] > profiles/..."""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from harness import workloads as W  # noqa: E402

level = sys.argv[1] if len(sys.argv) > 1 else "fused"
dev = torch.device("cuda", 0)
net = W.build(9, level, dev)
if "capturable" in sys.argv:      # the graph-mode rebindings of the harness (cached DySample grid, min pool without im2col)
    from harness import graph_step
    graph_step.make_capturable()
step = W.TrainStep(net, 9)
x, y = W.synthetic_batch(24, 224, 9)
for _ in range(3):
    step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(x, y)
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = ev.name[:90]
        tot[k][0] += 1
        tot[k][1] += ev.device_time
total = sum(v[1] for v in tot.values())
print(f"# level={level}: {sum(v[0] for v in tot.values())} GPU kernels/memcpys, {total / 1e3:.2f} ms of GPU time in one train step")
cats = defaultdict(float)
for k, (n, t) in tot.items():
    kl = k.lower()
    if "ss2d::" in k or "scan_" in kl or "out_gate" in kl or "wgrad_ts" in kl or "dwconv3_wgrad" in kl or "layernorm_fwd_kernel" in kl or "layernorm_bwd_kernel" in kl or "cross_" in kl:
        c = "ours (libss2d_b200)"
    elif "gemm" in kl or "cutlass" in kl or "xmma" in kl or "cublas" in kl or "gemv" in kl:
        c = "library GEMM"
    elif "conv" in kl or "cudnn" in kl or "wgrad" in kl or "dgrad" in kl:
        c = "library conv"
    elif "elementwise" in kl or "vectorized" in kl:
        c = "torch elementwise"
    elif "reduce" in kl:
        c = "torch reduce"
    elif "memcpy" in kl or "memset" in kl or "copy" in kl or "cat" in kl:
        c = "copies / cat / memset"
    elif "norm" in kl:
        c = "torch norm"
    else:
        c = "other"
    cats[c] += t
for c, t in sorted(cats.items(), key=lambda kv: -kv[1]):
    print(f"## {c:28s} {t / 1e3:8.2f} ms  {100 * t / total:5.1f}%")
print(f"{'kernel':92s} {'n':>5s} {'ms':>8s} {'%':>6s}")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{k:92s} {n:5d} {t / 1e3:8.3f} {100 * t / total:6.2f}")
