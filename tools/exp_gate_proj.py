import sys, statistics
sys.path.insert(0, "/root/repo")
import torch
from ceigm_unet_b200 import ops
Bn, D, H, W, C = 24, 192, 56, 56, 96
L = H * W
lnw, lnb = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts) * 1e3
for dtype in (torch.float32, torch.bfloat16):
    z = torch.randn(Bn, L, D, device="cuda").to(dtype)
    Wt = (torch.randn(C, D, device="cuda") / D ** 0.5).to(dtype)
    for K, tmask, name in ((4, 0b1010, "K4 nat+tr"), (4, 0b0000, "K4 all natural"), (4, 0b1111, "K4 all transposed"), (1, 0, "K1 natural"), (2, 0b10, "K2 nat+tr")):
        ys = torch.randn(Bn, K, D, L, device="cuda")
        a = t(lambda: ops.gate_proj_fwd(ys, lnw, lnb, z, True, 1e-5, Wt, None, (H, W), tmask, False))
        b = t(lambda: ops.gate_proj_fwd(ys, lnw, lnb, None, True, 1e-5, Wt, None, (H, W), tmask, False))
        c = t(lambda: ops.gate_proj_fwd(ys, lnw, lnb, z, True, 1e-5, Wt, None, (H, W), tmask, True))
        print(f"{str(dtype)[6:]:9s} {name:18s} with z {a:7.1f} us   no z {b:7.1f} us   with z + g_out {c:7.1f} us")
