"""Timing of the NATURAL-layout (fused cross-scan/merge) scan through ops.ScanProblem, per direction set.
    python tools/bench_natural.py [B] [D] [H] [N]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ceigm_unet_b200 import ops
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
D = int(sys.argv[2]) if len(sys.argv) > 2 else 192
H = int(sys.argv[3]) if len(sys.argv) > 3 else 56
N = int(sys.argv[4]) if len(sys.argv) > 4 else 16
W, L = H, H * H
dev = "cuda:0"
torch.manual_seed(0)
def t(*s): return torch.randn(*s, device=dev)
for dirs in [(1, 1, 1, 1), (3, 3, 3, 3), (2, 2, 2, 2), (4, 4, 4, 4), (1, 2, 3, 4)]:
    K = len(dirs)
    x, dts = t(Bn, D, L), 0.5 * torch.rand(Bn, K * D, L, device=dev)
    Bs, Cs = t(Bn, K, N, L), t(Bn, K, N, L)
    A, Dv, bias = -0.5 * torch.rand(K * D, N, device=dev), t(K * D), 0.5 * torch.rand(K * D, device=dev)
    dy = t(Bn, D, L)
    prob = ops.ScanProblem(x, dts, A, Bs, Cs, Dv, bias, True, out_float=True, hw=(H, W), dirs=dirs, u_mod=D)
    def run(bwd):
        out, st = prob.forward(True)
        if bwd: prob.backward(dy, st)
    res = {}
    for name, bwd in (("fwd", False), ("fwd+bwd", True)):
        g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        with torch.cuda.stream(s):
            run(bwd); torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s): run(bwd)
        torch.cuda.synchronize()
        for _ in range(2): g.replay()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[name] = statistics.median(ts)
    # level-2 algorithmic bytes: fwd s((2+K) B D L + 2 B K N L) with per-direction outputs counted K times
    print(f"dirs={dirs} B={Bn} D={D} L={L} N={N}: fwd {res['fwd']:.3f} ms  bwd {res['fwd+bwd']-res['fwd']:.3f} ms")
