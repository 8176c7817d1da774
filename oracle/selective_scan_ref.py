"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the reference's pure-PyTorch selective scan,
`selective_scan_ref` at /root/reference/gm-unet/kernels/selective_scan/test_selective_scan.py:168-234.

Restated, not copied: only the real-valued, input-dependent-B/C path that GM-UNet uses
(B, C of shape (batch, groups, dstate, L) or (batch, dstate, L)); the complex and constant-B/C
branches (`:193-200`, `:203-204`, `:218-219`, `:226-227`) are not on the hot path and are omitted.
The `Tensor.embed_dim()` typo at `:191-192` is read as `.dim()`.

Parity status: PINNED against the reference function itself, executed unmodified in the build
container by tests/golden/make_golden.py (fixtures tests/golden/scan_*.npz).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                       delta_softplus=False, return_last_state=False):
    """Sequential scan, differentiable by autograd (ref :183-234).

    u, delta: (b, d, L); A: (d, n) real; B, C: (b, n, L) or (b, g, n, L); D, delta_bias: (d,).
    Returns out (b, d, L) in u's dtype [, last_state (b, d, n) fp32].
    """
    in_dtype = u.dtype
    u32 = u.float()
    dt = delta.float()
    if delta_bias is not None:                       # ref :186-187
        dt = dt + delta_bias.float().unsqueeze(-1)
    if delta_softplus:                               # ref :188-189 (F.softplus: threshold 20)
        dt = F.softplus(dt)
    nb, nd, nstate = u.shape[0], A.shape[0], A.shape[1]
    Bf, Cf = B.float(), C.float()                    # ref :199-200
    if Bf.dim() == 4:                                # ref :209-211 grouped B -> per-channel
        Bf = Bf.repeat_interleave(nd // Bf.shape[1], dim=1)
        dBu = dt.unsqueeze(-1) * Bf.permute(0, 1, 3, 2) * u32.unsqueeze(-1)     # (b,d,L,n)
    else:                                            # ref :206-207
        dBu = dt.unsqueeze(-1) * Bf.permute(0, 2, 1).unsqueeze(1) * u32.unsqueeze(-1)
    if Cf.dim() == 4:                                # ref :212-213
        Cf = Cf.repeat_interleave(nd // Cf.shape[1], dim=1)
    dA = torch.exp(dt.unsqueeze(-1) * A.float().view(1, nd, 1, nstate))     # ref :203  (b,d,L,n)
    state = A.new_zeros((nb, nd, nstate), dtype=torch.float32)
    cols = []
    for i in range(u.shape[2]):                      # ref :215-229
        state = dA[:, :, i] * state + dBu[:, :, i]
        if Cf.dim() == 3:
            cols.append((state * Cf[:, :, i].unsqueeze(1)).sum(-1))
        else:
            cols.append((state * Cf[:, :, :, i]).sum(-1))
    y = torch.stack(cols, dim=2)
    out = y if D is None else y + u32 * D.float().unsqueeze(-1)       # ref :230
    if z is not None:                                # ref :231-232
        out = out * F.silu(z.float())
    out = out.to(in_dtype)                           # ref :233
    return (out, state) if return_last_state else out


def make_inputs(batch, dim, seqlen, dstate, groups=1, *, has_D=True, has_delta_bias=True,
                dtype=torch.float32, seed=0, device="cpu", requires_grad=False):
    """Seeded input recipe of the reference test (test_selective_scan.py:406-441):
    A=-0.5*rand(D,N); B,C=randn(B,G,N,L); D=randn(D); delta_bias=0.5*rand(D); u=randn(B,D,L);
    delta=0.5*rand(B,D,L). Generated on CPU with a private generator, then moved."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = (-0.5 * torch.rand(dim, dstate, generator=g)).float()
    Bm = torch.randn(batch, groups, dstate, seqlen, generator=g).to(dtype)
    Cm = torch.randn(batch, groups, dstate, seqlen, generator=g).to(dtype)
    Dv = torch.randn(dim, generator=g).float() if has_D else None
    bias = (0.5 * torch.rand(dim, generator=g)).float() if has_delta_bias else None
    u = torch.randn(batch, dim, seqlen, generator=g).to(dtype)
    delta = (0.5 * torch.rand(batch, dim, seqlen, generator=g)).to(dtype)
    dout = torch.randn(batch, dim, seqlen, generator=g).to(dtype)
    out = dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=Dv, delta_bias=bias, dout=dout)
    for k, v in out.items():
        if v is not None:
            v = v.to(device)
            if requires_grad and k != "dout":
                v.requires_grad_(True)
            out[k] = v
    return out
