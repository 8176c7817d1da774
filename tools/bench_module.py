"""fwd+bwd time of one SS2D module call: the fused path (modules.SS2D) vs the reference's op-by-op composition
(CrossScan -> einsum projections -> SelectiveScanCore -> CrossMerge -> transpose -> LayerNorm -> gate), both on our
kernels.   python tools/bench_module.py [d_model] [d_state] [ssm_ratio] [K] [B] [H]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import ceigm_unet_b200 as P
a = [float(v) for v in sys.argv[1:]]
d_model, d_state, ratio, K, Bn, H = (int(a[0]) if a else 96), (int(a[1]) if len(a) > 1 else 16), (a[2] if len(a) > 2 else 2.0), \
    (int(a[3]) if len(a) > 3 else 4), (int(a[4]) if len(a) > 4 else 24), (int(a[5]) if len(a) > 5 else 56)
torch.manual_seed(0)
m = P.SS2D(d_model=d_model, d_state=d_state, ssm_ratio=ratio, k_group=K).cuda()
x = torch.randn(Bn, H, H, d_model, device="cuda", requires_grad=True)
dirs = (1, 2, 3, 4) if K == 4 else (2,)
pair = (P.CrossScan, P.CrossMerge) if K == 4 else (P.CrossScan_2, P.CrossMerge_2)

def reference_style(mod, x):
    """ss2d.py:349-519 op by op (what level-1 drop-in gives)."""
    xz = mod.in_proj(x); xi, z = xz.chunk(2, dim=-1); z = F.silu(z)
    xi = F.silu(mod.conv2d(xi.permute(0, 3, 1, 2).contiguous()))
    Bb, D, Hh, Ww = xi.shape; L = Hh * Ww; Kk, _, R = mod.dt_projs_weight.shape; N = mod.A_logs.shape[1]
    xs = pair[0].apply(xi)
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, mod.x_proj_weight)
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = torch.einsum("bkrl,kdr->bkdl", dts, mod.dt_projs_weight)
    ys = P.SelectiveScanCore.apply(xs.view(Bb, -1, L), dts.contiguous().view(Bb, -1, L), -torch.exp(mod.A_logs.float()),
                                   Bs.contiguous(), Cs.contiguous(), mod.Ds.float(), mod.dt_projs_bias.view(-1).float(), True)
    y = pair[1].apply(ys.view(Bb, Kk, -1, Hh, Ww))
    y = y.view(Bb, -1, L).transpose(1, 2).contiguous().view(Bb, Hh, Ww, -1)
    y = mod.out_norm(y) * z
    return mod.out_proj(y)

def fused(mod, x):
    return mod(x) if K == 4 else mod(x, CrossScan=pair[0], CrossMerge=pair[1])

def time_it(fn, iters=10):
    def step():
        y = fn(m, x); y.backward(gy)
        x.grad = None
        for p_ in m.parameters(): p_.grad = None
    gy = torch.randn(Bn, H, H, d_model, device="cuda")
    for _ in range(3): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)

with torch.no_grad():
    y1, y2 = fused(m, x), reference_style(m, x)
print("max |fused - op-by-op| / max|ref| = %.2e" % ((y1 - y2).abs().max() / y2.abs().max()).item())
tf, tr = time_it(fused), time_it(reference_style)
print(f"SS2D d_model={d_model} N={d_state} D={int(ratio*d_model)} K={K} B={Bn} {H}x{H}: fused fwd+bwd {tf:.3f} ms | op-by-op on our kernels {tr:.3f} ms | x{tr/tf:.2f}")
