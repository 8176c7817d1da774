"""Development check of the fast-path scan kernels (scan_fwd2.cu / scan_bwd2.cu) against an fp64 torch restatement of
selective_scan_ref run on the GPU (sequential over L, autograd for the gradients). Covers SCAN layout, NATURAL layout
with directions 1 / 3, ragged channel counts, padded state counts and tails. Usage: python tools/dev_v2.py [--bwd]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import ceigm_unet_b200 as P
from ceigm_unet_b200 import ops

dev = "cuda:0"
do_bwd = "--bwd" in sys.argv


def ref(u, dl, A, B, C, D, bias, softplus, dirs=None):
    """fp64; tensors in NATURAL order when dirs is given (dir 1 = identity, 3 = reversed), else scan order."""
    b, dt, L = dl.shape
    G, N = B.shape[1], B.shape[2]
    dpg = dt // G
    if u.shape[1] != dt:
        u = u.repeat(1, dt // u.shape[1], 1)
    flip = torch.zeros(G, dtype=torch.bool)
    if dirs is not None:
        flip = torch.tensor([k == 3 for k in dirs])
    def f(x, per_group_of):     # flip the groups that run reversed
        if not flip.any():
            return x
        xs = list(x.split(per_group_of, dim=1))
        return torch.cat([t.flip(-1) if flip[i] else t for i, t in enumerate(xs)], dim=1)
    u_, dl_, B_, C_ = f(u, dpg), f(dl, dpg), f(B, 1), f(C, 1)
    x = dl_ + bias[None, :, None]
    d_ = torch.where(x > 20, x, torch.log1p(torch.exp(x))) if softplus else x
    Bx = B_.repeat_interleave(dpg, dim=1)      # (b, dt, N, L)
    Cx = C_.repeat_interleave(dpg, dim=1)
    h = torch.zeros(b, dt, N, dtype=u.dtype, device=u.device)
    ys = []
    for l in range(L):
        a = torch.exp(d_[:, :, l, None] * A[None])
        h = a * h + (d_[:, :, l] * u_[:, :, l])[..., None] * Bx[:, :, :, l]
        ys.append((h * Cx[:, :, :, l]).sum(-1))
    y = torch.stack(ys, dim=-1) + D[None, :, None] * u_
    return f(y, dpg), h


def case(name, b, dt, L, N, G, hw=None, dirs=None, u_mod=0, softplus=True, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    A = -0.5 * torch.rand(dt, N, device=dev, generator=g)
    B = torch.randn(b, G, N, L, device=dev, generator=g)
    C = torch.randn(b, G, N, L, device=dev, generator=g)
    D = torch.randn(dt, device=dev, generator=g)
    bias = 0.5 * torch.rand(dt, device=dev, generator=g)
    u = torch.randn(b, u_mod if u_mod else dt, L, device=dev, generator=g)
    dl = 0.5 * torch.rand(b, dt, L, device=dev, generator=g)
    if softplus:
        dl[0, 0, : min(8, L)] = 25.0      # exercise the threshold branch
    pr = ops.ScanProblem(u, dl, A, B, C, D, bias, softplus, hw=hw, dirs=dirs, u_mod=u_mod)
    out, x = pr.forward(True)
    t64 = [t.double().requires_grad_(True) for t in (u, dl, A, B, C, D, bias)]
    yref, hlast = ref(*t64, softplus, dirs)
    def rel(a, r):
        return float((a.double() - r).abs().max() / r.abs().max().clamp_min(1e-30))
    errs = {"out": rel(out, yref.detach()), "last": rel(x[:, :, -1, 1::2], hlast.detach())}
    if do_bwd:
        dout = torch.randn(b, u_mod if u_mod else dt, L, device=dev, generator=g)
        grads = pr.backward(dout, x)
        dy = dout.double()
        if u_mod:
            dy = dy.repeat(1, dt // u_mod, 1)
        (yref * dy).sum().backward()
        names = ["du", "ddelta", "dA", "dB", "dC", "dD", "dbias"]
        for nme, gg, t in zip(names, grads, t64):
            r = t.grad
            if nme == "du" and u_mod:
                # ours: one plane per direction (B, dt, L); reference: summed over the repeats
                gg = gg.view(b, dt // u_mod, u_mod, L).sum(1)
            errs[nme] = rel(gg, r)
    worst = max(errs.values())
    print(("OK  " if worst < 1e-3 else "FAIL"), name, {k: "%.1e" % v for k, v in errs.items()}, flush=True)
    return worst < 1e-3


ok = True
ok &= case("scan B2 Dt64 L256 N16 G4", 2, 64, 256, 16, 4)
ok &= case("scan B1 Dt160 L100 N16 G4 (ragged rows, tail)", 1, 160, 100, 16, 4)
ok &= case("scan B2 Dt32 L64 N12 G2 (padded states)", 2, 32, 64, 12, 2)
ok &= case("scan B1 Dt192 L512 N16 G1 nosoftplus", 1, 192, 512, 16, 1, softplus=False)
ok &= case("natural dirs 1,3,1,3 B2 Dt96 8x12 N16", 2, 96, 96, 16, 4, hw=(8, 12), dirs=[1, 3, 1, 3])
ok &= case("natural dirs 1,3 shared u B2 D24 10x10 N16", 2, 48, 100, 16, 2, hw=(10, 10), dirs=[1, 3], u_mod=24)
ok &= case("natural dir 3 long B1 Dt40 L1000 N16", 1, 40, 1000, 16, 1, hw=(25, 40), dirs=[3])
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
