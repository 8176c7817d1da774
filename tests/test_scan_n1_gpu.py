"""GPU parity tests of the lean d_state = 1 kernels (csrc/scan_n1.cu: forward, one-row backward, rows-walking backward with
the TMA-staged ring) through the C ABI against the C/f64 oracle, at the live GM-UNet shapes (16 / 32 / 87 / 112
rows per direction on 56^2 ... 7^2 maps, alone and as the four grouped SS2Ds of a GroupMambaLayer) and at the geometry
corners: sub-warp rows (L = 20, 49), one warp per row (L = 132, 196), several warps per row (L = 784, 3136), multi-chunk rows
(L = 4100 > 16 warps x 256), unaligned lengths (L % 4 != 0: scalar path), reversed traversals, the input shared between
groups (u_dim_modulo), row counts that leave row lanes and row blocks ragged, no softplus / no D / no bias.
Tolerance: rel <= 1e-3 fp32 (max |got - ref| / max |ref| per tensor), as in tests/test_scan_fast_gpu.py."""
import numpy as np
import pytest
import torch

from test_scan_fast_gpu import _case

pytestmark = pytest.mark.gpu

LIVE = [(16, 56), (32, 28), (87, 14), (112, 7)]           # (rows per direction, H = W)


@pytest.mark.parametrize("D,H", LIVE, ids=[f"D{d}_{h}x{h}" for d, h in LIVE])
def test_live_shapes_single_direction_scan_layout(D, H):
    _case(3, D, H * H, 1, 1, seed=D)


@pytest.mark.parametrize("D,H", LIVE, ids=[f"D{d}_{h}x{h}" for d, h in LIVE])
def test_live_shapes_grouped_four_directions(D, H):
    """One launch for the four SS2Ds of a layer: groups 2 and 4 are row-major / reversed traversals of transposed planes."""
    _case(2, 4 * D, H * H, 1, 4, hw=(H, H), dirs=[1, 3, 1, 3], seed=D + 1)


@pytest.mark.parametrize("L", [20, 49, 132, 196, 260, 784, 1000, 3136, 4096, 4100, 9000])
def test_geometry_corners(L):
    _case(2, 10, L, 1, 2, seed=L)


@pytest.mark.parametrize("L,hw", [(49, (7, 7)), (196, (14, 14)), (788, (4, 197)), (3136, (56, 56)), (4100, (50, 82))])
def test_reversed_and_shared_input(L, hw):
    _case(2, 3 * 7, L, 1, 3, hw=hw, dirs=[3, 1, 3], u_mod=7, seed=L + 3)


@pytest.mark.parametrize("rows", [1, 2, 5, 17, 40, 87, 200])
def test_ragged_row_counts(rows):
    _case(1, 2 * rows, 256, 1, 2, seed=rows)              # rows kernel: partial row lanes and partial row blocks
    _case(5, rows, 520, 1, 1, seed=rows + 1)


def test_flags_off():
    _case(2, 24, 784, 1, 1, softplus=False, has_D=False, has_bias=False)
    _case(2, 24, 50, 1, 2, softplus=False, has_D=True, has_bias=False)


def test_unaligned_views_take_the_scalar_path():
    """Row starts that are not 16-byte aligned (a slice along L) must not reach the vector / TMA kernels."""
    from ceigm_unet_b200 import ops
    from oracle import c_oracle
    torch.manual_seed(0)
    b, dt, L = 2, 12, 257
    big = lambda *s: torch.randn(*s, device="cuda")
    u, dl, dout = big(b, dt, L + 3)[..., 1:1 + L], 0.5 * torch.rand(b, dt, L + 3, device="cuda")[..., 1:1 + L], big(b, dt, L)
    A, D, bias = -0.5 * torch.rand(dt, 1, device="cuda"), big(dt), 0.5 * torch.rand(dt, device="cuda")
    B, C = big(b, 1, 1, L + 1)[..., 1:], big(b, 1, 1, L + 1)[..., 1:]
    pr = ops.ScanProblem(u, dl, A, B, C, D, bias, True)
    out, x = pr.forward(True)
    grads = pr.backward(dout, x)
    n = lambda t: t.float().cpu().numpy()
    ref_out, _ = c_oracle.scan_fwd(n(u), n(dl), n(A), n(B), n(C), n(D), n(bias), True)
    ref = c_oracle.scan_bwd(n(u), n(dl), n(A), n(B), n(C), n(D), n(bias), n(dout), True)
    rel = lambda a, r: float(np.abs(n(a).astype(np.float64) - r).max() / max(np.abs(r).max(), 1e-30))
    assert rel(out, ref_out) < 1e-3
    for name, g in zip(["du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias"], grads):
        assert rel(g, ref[name]) < 1e-3, name


def test_backward_is_deterministic_when_one_cta_covers_a_group():
    """Small groups are walked by a single CTA: dB / dC are plain stores of sums taken in a fixed order."""
    from ceigm_unet_b200 import ops
    from oracle.selective_scan_ref import make_inputs
    a = make_inputs(64, 64, 784, 1, groups=4, seed=5, device="cuda")       # 256 (batch, group) pairs >= 1.5 x 148: 1 CTA each
    pr = ops.ScanProblem(a["u"], a["delta"], a["A"], a["B"], a["C"], a["D"], a["delta_bias"], True)
    _, x = pr.forward(True)
    g1 = pr.backward(a["dout"], x)
    g2 = pr.backward(a["dout"], x)
    assert torch.equal(g1[3], g2[3]) and torch.equal(g1[4], g2[4])
