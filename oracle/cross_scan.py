"""ORACLE (test infrastructure only — never imported by the product path).

Restatement of the reference's cross-scan / cross-merge permutations as explicit index maps.
Follows /root/reference/gm-unet/model/gm/csms6s.py:11-206 (forward definitions). Backward here is
the TRUE adjoint (autograd of the forward); the reference's hand-written `CrossScan_2/_4.backward`
(`:108`, `:185`) equal it only for H == W — every config is square.

Direction numbering (1-based as in the reference class names; K=4 `CrossScan` stacks them in this order):
  1: row-major (h, w)        xs[l] = x[h, w],  l = h*W + w
  2: column-major            xs[l] = x[h, w],  l = w*H + h
  3: reversed row-major      xs[l] = x[h, w],  l = L-1 - (h*W + w)
  4: reversed column-major   xs[l] = x[h, w],  l = L-1 - (w*H + h)

Parity status: PINNED against the reference classes executed unmodified (tests/golden/cross_*.npz).
"""
from __future__ import annotations

import torch


def scan_index(direction: int, H: int, W: int) -> torch.Tensor:
    """idx[l] = flat natural offset (h*W + w) of the pixel visited at scan position l."""
    nat = torch.arange(H * W).view(H, W)
    if direction in (1, 3):
        idx = nat.reshape(-1)
    elif direction in (2, 4):
        idx = nat.t().reshape(-1)
    else:
        raise ValueError(direction)
    if direction in (3, 4):
        idx = idx.flip(0)
    return idx


def cross_scan_k(x: torch.Tensor, direction: int) -> torch.Tensor:
    """CrossScan_<direction>.forward: (B, C, H, W) -> (B, 1, C, L). ref csms6s.py:57-64, 95-102, 134-141, 172-179."""
    Bn, Cn, H, W = x.shape
    idx = scan_index(direction, H, W).to(x.device)
    return x.reshape(Bn, Cn, H * W).index_select(-1, idx).unsqueeze(1)


def cross_merge_k(ys: torch.Tensor, direction: int) -> torch.Tensor:
    """CrossMerge_<direction>.forward: (B, 1, D, H, W)-shaped scan-order data -> (B, D, L) natural order.
    ref csms6s.py:75-81, 112-118, 153-159, 190-196."""
    Bn, K, D, H, W = ys.shape
    assert K == 1
    idx = scan_index(direction, H, W).to(ys.device)
    out = torch.empty(Bn, D, H * W, dtype=ys.dtype, device=ys.device)
    out[..., idx] = ys.reshape(Bn, D, H * W)
    return out


def cross_scan4(x: torch.Tensor) -> torch.Tensor:
    """CrossScan.forward (K=4): (B, C, H, W) -> (B, 4, C, L). ref csms6s.py:12-20."""
    return torch.cat([cross_scan_k(x, k) for k in (1, 2, 3, 4)], dim=1)


def cross_merge4(ys: torch.Tensor) -> torch.Tensor:
    """CrossMerge.forward (K=4): (B, 4, D, H, W) -> (B, D, L); summation order of ref csms6s.py:38-39:
    (ys0 + unflip(ys2)) + transpose(ys1 + unflip(ys3))."""
    Bn, K, D, H, W = ys.shape
    assert K == 4
    m = [cross_merge_k(ys[:, k:k + 1], k + 1) for k in range(4)]
    # m[k] are all in natural order; (m0 + m2) + (m1 + m3) reproduces the reference's association
    return (m[0] + m[2]) + (m[1] + m[3])
