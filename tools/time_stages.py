#!/usr/bin/env python
"""Per-stage timing of the live (d_state = 1) scan calls, ours vs the reference CUDA kernel, CUDA-graph replay.
python tools/time_stages.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ceigm_unet_b200.dropin import selective_scan_cuda_core as core  # noqa: E402
from harness import refmodel  # noqa: E402

dev = torch.device("cuda:0")
ref = refmodel.load_ref_cuda_ext()
names = sys.argv[1:] or ["gm_s1", "gm_s2", "gm_s3", "gm_s4", "gm_s1_g4", "gm_s2_g4", "gm_s3_g4", "gm_s4_g4"]
print(f"{'workload':12s} {'MB fwd':>8s} {'ours fwd us':>12s} {'ours f+b us':>12s} {'ref fwd us':>11s} {'ref f+b us':>11s}")
for n in names:
    calls = bench.WORKLOADS[n]
    sets = bench.build_inputs(calls, dev, on_device=True)
    fb, bb = bench.alg_bytes(calls)
    r = [bench.graph_time_ms(core, sets, with_bwd=False, iters=20), bench.graph_time_ms(core, sets, with_bwd=True, iters=20)]
    if ref is not None:
        r += [bench.graph_time_ms(ref, sets, with_bwd=False, iters=20), bench.graph_time_ms(ref, sets, with_bwd=True, iters=20)]
    else:
        r += [float("nan")] * 2
    print(f"{n:12s} {fb / 1e6:8.2f} " + " ".join(f"{v * 1e3:11.2f}" for v in r))
