"""ORACLE (test infrastructure only — never imported by the product path).

Functional restatement (plain torch ops, autograd-differentiable, CPU) of the composites around
the scan, written as functions of a `state_dict`-style mapping so that the product modules can be
checked by copying weights:

  ss2d_core      : SS2D.forward_corev2    /root/reference/gm-unet/model/gm/ss2d.py:459-500
                                          (≡ model/vmamba/vmamba.py:639-685)
  ss2d_forward   : SS2D.forwardv2         model/gm/ss2d.py:502-519 (≡ model/vmamba/vmamba.py:687-720)
  group_layer    : GroupMambaLayer.forward model/gm/groupmamba.py:127-159

`directions` is the tuple of 1-based scan directions, one per k (gm: a single direction (k,);
vm: (1, 2, 3, 4)). Parity status: PINNED against the reference modules executed unmodified
(tests/golden/ss2d_*.npz, group_mamba_layer.npz).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .cross_scan import cross_merge_k, cross_scan_k
from .selective_scan_ref import selective_scan_ref

# Module-level checks at the live shapes swap in the C-backed scan (oracle/fast_scan.py: same function, analytic backward);
# the default stays the restated reference loop so that the golden-pinned tests exercise it.
_SCAN = [selective_scan_ref]


def use_fast_scan(on: bool = True) -> None:
    from .fast_scan import selective_scan_fast
    _SCAN[0] = selective_scan_fast if on else selective_scan_ref


def ss2d_core(x, p, directions, force_fp32=True):
    """x: (B, D, H, W) -> (B, H, W, D) after out_norm. p: mapping with x_proj_weight (K,R+2N,D),
    dt_projs_weight (K,D,R), dt_projs_bias (K,D), A_logs (K*D,N), Ds (K*D), out_norm.weight/bias."""
    Bn, D, H, W = x.shape
    K, _, R = p["dt_projs_weight"].shape
    N = p["A_logs"].shape[1]
    L = H * W
    xs = torch.cat([cross_scan_k(x, k) for k in directions], dim=1)              # ss2d.py:459 (B,K,D,L)
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, p["x_proj_weight"])                # :465
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)                            # :468
    dts = torch.einsum("bkrl,kdr->bkdl", dts, p["dt_projs_weight"])               # :469
    u = xs.reshape(Bn, K * D, L)
    dts = dts.contiguous().view(Bn, K * D, L)
    As = -torch.exp(p["A_logs"].float())                                          # :473
    Ds = p["Ds"].float()
    bias = p["dt_projs_bias"].reshape(-1).float()
    if force_fp32:                                                                # :479-480
        u, dts, Bs, Cs = u.float(), dts.float(), Bs.float(), Cs.float()
    ys = _SCAN[0](u, dts, As, Bs.contiguous(), Cs.contiguous(), Ds, None, bias, True)
    ys = ys.view(Bn, K, D, H, W)                                                  # :484
    parts = [cross_merge_k(ys[:, i:i + 1], k) for i, k in enumerate(directions)]  # :486
    if len(parts) == 4:
        y = (parts[0] + parts[2]) + (parts[1] + parts[3])                         # csms6s.py:38-39
    else:
        y = parts[0]
        for q in parts[1:]:
            y = y + q
    y = y.view(Bn, D, L).transpose(1, 2).contiguous().view(Bn, H, W, D)           # :495-497
    y = F.layer_norm(y, (D,), p["out_norm.weight"], p["out_norm.bias"], 1e-5)     # :498
    return y.to(x.dtype)


def ss2d_forward(x, p, directions):
    """x: (B, H, W, C) -> (B, H, W, C). p additionally has in_proj.weight (2D,C), conv2d.weight (D,1,k,k),
    conv2d.bias (D), out_proj.weight (C,D)."""
    xz = F.linear(x, p["in_proj.weight"], p.get("in_proj.bias"))                  # :504
    xi, z = xz.chunk(2, dim=-1)                                                   # :506
    z = F.silu(z)                                                                 # :508
    xi = xi.permute(0, 3, 1, 2).contiguous()                                      # :510
    D = xi.shape[1]
    kk = p["conv2d.weight"].shape[-1]
    xi = F.conv2d(xi, p["conv2d.weight"], p.get("conv2d.bias"), padding=(kk - 1) // 2, groups=D)   # :512
    xi = F.silu(xi)                                                               # :513
    y = ss2d_core(xi, p, directions)                                              # :514
    y = y * z                                                                     # :517
    return F.linear(y, p["out_proj.weight"], p.get("out_proj.bias"))              # :518


def _sub(p, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


def group_layer(x, p, H, W):
    """GroupMambaLayer.forward: x (B, L, C) -> (B, L, C_out). groupmamba.py:127-159."""
    Bn, L, C = x.shape
    xn = F.layer_norm(x, (C,), p["norm.weight"], p["norm.bias"], 1e-5)            # :131
    zmean = xn.permute(0, 2, 1).mean(dim=2)                                       # :134
    a1 = F.relu(F.linear(zmean, p["fc1.weight"], p["fc1.bias"]))                  # :136
    aff = torch.sigmoid(F.linear(a1, p["fc2.weight"], p["fc2.bias"]))             # :137
    x4 = xn.view(Bn, H, W, C)
    chunks = torch.chunk(x4, 4, dim=-1)                                           # :140
    outs = [ss2d_forward(chunks[i], _sub(p, f"mamba_g{i + 1}."), (i + 1,)) for i in range(4)]   # :143-146
    xm = torch.cat(outs, dim=-1) * p["skip_scale"] * x4                           # :149
    xm = xm.view(Bn, L, C) * aff.unsqueeze(1)                                     # :151-154
    xm = F.layer_norm(xm, (C,), p["norm.weight"], p["norm.bias"], 1e-5)           # :156 (same norm, shared weights)
    return F.linear(xm, p["proj.weight"], p["proj.bias"])                         # :157
