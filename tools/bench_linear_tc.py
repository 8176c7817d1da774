"""Device timing of the tensor-core projection kernel against the library composition it replaces
(F.linear + chunk + permute().contiguous(); F.linear), CUDA-graph replay: python tools/bench_linear_tc.py"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
for _a in sys.argv[1:]:
    if _a.startswith('--lib='):
        import ceigm_unet_b200  # noqa
        sys.modules['ceigm_unet_b200._lib'].LIB_PATH = os.path.abspath(_a[6:])
from ceigm_unet_b200 import ops

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.synchronize()
    ts = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts) * 1e3

torch.backends.cuda.matmul.allow_tf32 = False
for dtype in (torch.float32, torch.bfloat16):
    for (Bn, H, W, C, D) in [(24, 56, 56, 96, 192), (24, 28, 28, 192, 384), (24, 56, 56, 64, 64)]:
        es = 4 if dtype == torch.float32 else 2
        x = torch.randn(Bn, H, W, C, device="cuda").to(dtype)
        Win = (torch.randn(2 * D, C, device="cuda") / C ** 0.5).to(dtype)
        Wout = (torch.randn(C, D, device="cuda") / D ** 0.5).to(dtype)
        y = torch.randn(Bn, H, W, D, device="cuda").to(dtype)
        M = Bn * H * W
        def lib_in():
            xz = F.linear(x, Win)
            xi, z = xz.chunk(2, dim=-1)
            return xi.permute(0, 3, 1, 2).contiguous(), z
        def tc_in():
            return ops.linear_tc(x, Win, None, [(D, ("planes", H * W), False), (D, "rows", False)])
        def lib_out(): return F.linear(y, Wout)
        def tc_out(): return ops.linear_tc(y, Wout, None, [(C, "rows", False)])
        ok_in = ops.linear_tc_supported(D, C, dtype); ok_out = ops.linear_tc_supported(C, D, dtype)
        b_in = es * M * (C + 2 * D); b_out = es * M * (C + D)
        t1 = timeit(lib_in); t2 = timeit(tc_in) if ok_in else float("nan")
        t3 = timeit(lib_out); t4 = timeit(tc_out) if ok_out else float("nan")
        for allow in (True,):
            pass
        print(f"{str(dtype)[6:]:9s} B{Bn} {H}x{W} C{C} D{D}: in_proj lib {t1:7.1f} us  tc {t2:7.1f} us ({b_in/t2/1e3:6.0f} GB/s alg)   "
              f"out_proj lib {t3:7.1f} us  tc {t4:7.1f} us ({b_out/t4/1e3:6.0f} GB/s alg)")
torch.backends.cuda.matmul.allow_tf32 = True
x = torch.randn(24, 56, 56, 96, device="cuda"); Win = torch.randn(384, 96, device="cuda") / 10
print("fp32 lib in_proj GEMM alone with TF32 allowed: %.1f us" % timeit(lambda: F.linear(x, Win)))
