"""Minimal driver for ncu captures: a few fwd+bwd scan calls of one bench workload, nothing else.
    python tools/prof_scan.py [workload] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from ceigm_unet_b200.dropin import selective_scan_cuda_core as core  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
sets = [(c, {k: v.to(dev) for k, v in inp.items()}) for c, inp in bench.build_inputs(bench.WORKLOADS[wl], dev)]
for _ in range(iters):
    for count, t in sets:
        out, x = core.fwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["delta_bias"], True, 1)
        core.bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["delta_bias"], t["dout"], x, True, 1)
torch.cuda.synchronize()
print("ok", wl, iters)
