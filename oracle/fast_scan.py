"""ORACLE (test infrastructure only — never imported by the product path).

`selective_scan_fast`: the same function as `selective_scan_ref` (oracle/selective_scan_ref.py, following
/root/reference/gm-unet/kernels/selective_scan/test_selective_scan.py:168-234) as a torch.autograd.Function whose
forward and backward run oracle/scan_oracle.c (double accumulation, OpenMP over rows; equations of
cus/selective_scan_bwd_kernel.cuh:125-272). Autograd through the Python loop of `selective_scan_ref` moves O(L^2) memory
(136 s for one K=4 call at L=3136), so module-level checks at the live shapes use this instead.

Parity status: PINNED — the C oracle is checked against the golden outputs and gradients of the reference's own
`selective_scan_ref` (tests/test_oracle_golden.py), and tests/test_oracle_golden.py::test_fast_scan_matches_ref checks this
wrapper against autograd of the restated reference on small shapes.
"""
from __future__ import annotations

import torch

from . import c_oracle


class _FastScan(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, delta, A, B, C, D, delta_bias, delta_softplus):
        squeeze = B.dim() == 3
        B4, C4 = (B.unsqueeze(1), C.unsqueeze(1)) if squeeze else (B, C)
        n = [None if t is None else t.detach().float().contiguous().numpy() for t in (u, delta, A, B4, C4, D, delta_bias)]
        out, _ = c_oracle.scan_fwd(*n, bool(delta_softplus), acc="f64")
        ctx.save_for_backward(u, delta, A, B, C, D if D is not None else u.new_empty(0),
                              delta_bias if delta_bias is not None else u.new_empty(0))
        ctx.meta = (squeeze, D is not None, delta_bias is not None, bool(delta_softplus))
        return torch.from_numpy(out).to(u.dtype)

    @staticmethod
    def backward(ctx, dout):
        u, delta, A, B, C, D, bias = ctx.saved_tensors
        squeeze, has_D, has_b, sp = ctx.meta
        B4, C4 = (B.unsqueeze(1), C.unsqueeze(1)) if squeeze else (B, C)
        f = lambda t: t.detach().float().contiguous().numpy()      # noqa: E731
        g = c_oracle.scan_bwd(f(u), f(delta), f(A), f(B4), f(C4), f(D) if has_D else None, f(bias) if has_b else None,
                              f(dout), sp, acc="f64")
        t = lambda a, like: torch.from_numpy(a).to(like.dtype)     # noqa: E731
        dB, dC = t(g["dB"], B), t(g["dC"], C)
        if squeeze:
            dB, dC = dB[:, 0], dC[:, 0]
        return (t(g["du"], u), t(g["ddelta"], delta), t(g["dA"], A), dB, dC,
                t(g["dD"], D) if has_D else None, t(g["ddelta_bias"], bias) if has_b else None, None)


def selective_scan_fast(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False):
    assert z is None, "the z gate is not on GM-UNet's path"
    return _FastScan.apply(u, delta, A, B, C, D, delta_bias, delta_softplus)
