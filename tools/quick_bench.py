"""Quick device-only timing of fwd / bwd for one workload: python tools/quick_bench.py [workload] [iters] [--graph]
--graph captures one step (all calls, fwd+bwd) in a CUDA graph and replays it: GPU time without host overhead."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, bench
for _a in sys.argv[1:]:
    if _a.startswith("--lib="):          # development A/B: time another build of the library (e.g. variants/base.so)
        import ceigm_unet_b200  # noqa: F401
        sys.modules["ceigm_unet_b200._lib"].LIB_PATH = os.path.abspath(_a[6:])
from ceigm_unet_b200.dropin import selective_scan_cuda_core as core
args = [a for a in sys.argv[1:] if not a.startswith("--")]
use_graph = "--graph" in sys.argv
wl = args[0] if len(args) > 0 else bench.DEFAULT_WORKLOAD
iters = int(args[1]) if len(args) > 1 else 20
dev = torch.device("cuda:0")
sets = [(c, {k: v.to(dev) for k, v in inp.items()}) for c, inp in bench.build_inputs(bench.WORKLOADS[wl], dev)]
fb, bb = bench.alg_bytes(bench.WORKLOADS[wl])
def run(rec=None, do_bwd=True):
    for count, t in sets:
        for _ in range(count):
            out, x = core.fwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["delta_bias"], True, 1)
            if rec: rec[0].record()
            if do_bwd:
                core.bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["delta_bias"], t["dout"], x, True, 1)
for _ in range(3): run()
torch.cuda.synchronize()
env = {k: v for k, v in os.environ.items() if k.startswith('SS2D_')}
if use_graph:
    res = {}
    for name, bwd in (("fwd", False), ("fwd+bwd", True)):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            run(do_bwd=bwd)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                run(do_bwd=bwd)
        torch.cuda.synchronize()
        for _ in range(3): g.replay()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name] = statistics.median(ts)
    f, fbk = res["fwd"], res["fwd+bwd"]
    print(f"{wl} GRAPH env={env} fwd {f:.4f} ms ({fb/f/1e6:.0f} GB/s)  bwd {fbk-f:.4f} ms ({bb/max(fbk-f,1e-9)/1e6:.0f} GB/s)  "
          f"fwd+bwd {fbk:.4f} ms ({(fb+bb)/fbk/1e6:.0f} GB/s)")
else:
    f, b = [], []
    for _ in range(iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[1].record(); run([e[0]]); e[2].record(); torch.cuda.synchronize()
        if len(sets) == 1 and sets[0][0] == 1:
            f.append(e[1].elapsed_time(e[0])); b.append(e[0].elapsed_time(e[2]))
        else:
            f.append(e[1].elapsed_time(e[2])); b.append(0.0)
    fm, bm = statistics.median(f), statistics.median(b)
    print(f"{wl} env={env} fwd {fm:.4f} ms ({fb/fm/1e6:.0f} GB/s)  bwd {bm:.4f} ms ({bb/max(bm,1e-9)/1e6:.0f} GB/s)")
