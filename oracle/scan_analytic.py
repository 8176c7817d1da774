"""ORACLE (test infrastructure only — never imported by the product path).

Fast restatement of the selective-scan forward and its hand-derived backward, so parity can be
checked at L=3136 where autograd through `selective_scan_ref` (O(L^2) memory traffic) is
impractical. Equations follow the reference CUDA backward,
/root/reference/gm-unet/kernels/selective_scan/csrc/selective_scan/cus/selective_scan_bwd_kernel.cuh:125-272
(SURVEY.md §8-a7), re-derived rather than transcribed:

    dt_l   = softplus(delta_l + bias_d)            (identity above 20)
    a_ln   = exp(dt_l * A_dn)
    h_ln   = a_ln * h_(l-1)n + dt_l * u_l * B_gnl ,  h_(-1) = 0
    y_l    = sum_n C_gnl * h_ln + D_d * u_l
    g_ln   = C_gnl * dy_l + a_(l+1)n * g_(l+1)n      (reverse scan)
    dC_gnl = sum_{d in g} dy_l * h_ln ;  dB_gnl = sum_{d in g} g_ln * dt_l * u_l
    du_l   = D_d * dy_l + dt_l * sum_n g_ln * B_gnl
    ddt_l  = u_l * sum_n g_ln B_gnl + sum_n g_ln * A_dn * a_ln * h_(l-1)n
    dA_dn  = sum_{b,l} g_ln * dt_l * a_ln * h_(l-1)n ;  dD_d = sum_{b,l} dy_l u_l
    ddelta = ddt * sigmoid(delta + bias) (softplus branch) ; dbias_d = sum_{b,l} ddelta

Parity status: PINNED — tests/test_oracle_golden.py checks this file against the golden outputs of
the reference's own `selective_scan_ref` + autograd (tests/golden/scan_*.npz).
"""
from __future__ import annotations

import torch


def _expand_groups(M, dim):
    # (b, g, n, L) or (b, n, L) -> (b, g, n, L)
    return M.unsqueeze(1) if M.dim() == 3 else M


@torch.no_grad()
def scan_fwd_bwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, dout=None,
                 acc_dtype=torch.float64):
    """Returns dict(out, last_state[, du, ddelta, dA, dB, dC, dD, ddelta_bias]) in `acc_dtype`."""
    T = acc_dtype
    squeeze_B, squeeze_C = B.dim() == 3, C.dim() == 3
    u_, dl = u.to(T), delta.to(T)
    A_ = A.to(T)
    Bg, Cg = _expand_groups(B.to(T), 0), _expand_groups(C.to(T), 0)
    nb, nd, L = u_.shape
    G, N = Bg.shape[1], Bg.shape[2]
    dpg = nd // G
    raw = dl + (delta_bias.to(T).view(1, nd, 1) if delta_bias is not None else 0.0)
    if delta_softplus:
        dt = torch.where(raw > 20.0, raw, torch.log1p(torch.exp(torch.clamp(raw, max=20.0))))
    else:
        dt = raw
    # per-channel views of B, C: (b, G, dpg, N, L) broadcast over dpg
    u5, dt5 = u_.view(nb, G, dpg, 1, L), dt.view(nb, G, dpg, 1, L)
    A5 = A_.view(1, G, dpg, N, 1)
    B5, C5 = Bg.view(nb, G, 1, N, L), Cg.view(nb, G, 1, N, L)
    h = torch.zeros(nb, G, dpg, N, dtype=T, device=u.device)
    keep_h = dout is not None
    hs = torch.empty(nb, G, dpg, N, L, dtype=T, device=u.device) if keep_h else None
    y = torch.empty(nb, G, dpg, L, dtype=T, device=u.device)
    for l in range(L):
        a = torch.exp(dt5[..., l] * A5[..., 0])
        h = a * h + (dt5[..., l] * u5[..., l]) * B5[..., l]
        if keep_h:
            hs[..., l] = h
        y[..., l] = (h * C5[..., l]).sum(-1)
    out = y.view(nb, nd, L)
    if D is not None:
        out = out + u_ * D.to(T).view(1, nd, 1)
    res = dict(out=out, last_state=h.reshape(nb, nd, N))
    if dout is None:
        return res
    dy = dout.to(T)
    dy5 = dy.view(nb, G, dpg, 1, L)
    g = torch.zeros_like(h)
    a_next = torch.zeros_like(h)
    du = torch.empty_like(u_).view(nb, G, dpg, L)
    ddt = torch.empty_like(du)
    dB = torch.zeros(nb, G, N, L, dtype=T, device=u.device)
    dC = torch.zeros_like(dB)
    dA = torch.zeros(G, dpg, N, dtype=T, device=u.device)
    for l in range(L - 1, -1, -1):
        a = torch.exp(dt5[..., l] * A5[..., 0])
        g = C5[..., l] * dy5[..., l] + a_next * g
        h_prev = hs[..., l - 1] if l > 0 else torch.zeros_like(g)
        dC[..., l] = (dy5[..., l] * hs[..., l]).sum(2)
        dB[..., l] = (g * (dt5[..., l] * u5[..., l])).sum(2)
        sB = (g * B5[..., l]).sum(-1)
        w = g * a * h_prev
        du[..., l] = dt5[..., 0, l] * sB
        ddt[..., l] = u5[..., 0, l] * sB + (w * A5[..., 0]).sum(-1)
        dA += (w * dt5[..., l]).sum(0)
        a_next = a
    du = du.view(nb, nd, L)
    ddt = ddt.view(nb, nd, L)
    if D is not None:
        du = du + dy * D.to(T).view(1, nd, 1)
        res["dD"] = (dy * u_).sum((0, 2))
    else:
        res["dD"] = None
    if delta_softplus:
        ddelta = torch.where(raw > 20.0, ddt, ddt * torch.sigmoid(raw))
    else:
        ddelta = ddt
    res.update(du=du, ddelta=ddelta, dA=dA.view(nd, N),
               dB=dB[:, 0] if squeeze_B else dB, dC=dC[:, 0] if squeeze_C else dC,
               ddelta_bias=ddelta.sum((0, 2)) if delta_bias is not None else None)
    return res
