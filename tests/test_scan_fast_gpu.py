"""GPU parity tests aimed at the fast paths of the north-star regime (fp32, TMA-stageable, 8 < d_state <= 16, SCAN
layout or directions 1 / 3) — the backward csrc/scan_bwd2.cu and the forwards that feed it its checkpoints: csrc/scan_fwd.cu
(8-row warps) and the lane-owns-row kernel csrc/scan_fwdr.cu, which only large calls select by themselves, so every case
runs once per forward kernel through the test hook ss2d_test_force_path (policy 1: 32-row warps, 2: 16-row warps,
3: 8-row warps, 4: the SEGMENTED mode of small calls — local scans per sequence segment + carry + fix-up kernels): idle warps, ragged channel counts, padded
state counts, sequence tails, reversed traversal with a tail (negative TMA start coordinate), input shared between
the groups (u_dim_modulo). Checked through the C ABI (ops.ScanProblem) against the C/f64 oracle; for direction 3 the
oracle sees the flipped sequences (CrossScan_3 / CrossMerge_3, model/gm/csms6s.py:133-168).

Tolerance: rel <= 1e-3 (fp32), max|got - ref| / max|ref| per tensor."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES = ["du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias"]


@pytest.fixture(params=[1, 2, 3, 4], ids=lambda p: "path%d" % p, autouse=True)
def forced_path(request):
    from ceigm_unet_b200 import _lib
    _lib.test_force_path(request.param)
    yield request.param
    _lib.test_force_path(0)


def _flip_groups(x, per_group, flip):
    """Flip the last axis of the channel groups that are traversed in reverse."""
    if not any(flip):
        return x
    parts = np.split(x, x.shape[1] // per_group, axis=1)
    return np.concatenate([p[..., ::-1] if flip[i] else p for i, p in enumerate(parts)], axis=1)


def _case(b, dt, L, N, G, hw=None, dirs=None, u_mod=0, softplus=True, seed=0, has_D=True, has_bias=True, a_scale=0.5):
    from ceigm_unet_b200 import ops
    from oracle import c_oracle
    gen = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    A = -a_scale * torch.rand(dt, N, device=dev, generator=gen)
    B = torch.randn(b, G, N, L, device=dev, generator=gen)
    C = torch.randn(b, G, N, L, device=dev, generator=gen)
    D = torch.randn(dt, device=dev, generator=gen) if has_D else None
    bias = 0.5 * torch.rand(dt, device=dev, generator=gen) if has_bias else None
    uch = u_mod if u_mod else dt
    u = torch.randn(b, uch, L, device=dev, generator=gen)
    dl = 0.5 * torch.rand(b, dt, L, device=dev, generator=gen)
    if softplus:
        dl[0, 0, : min(8, L)] = 25.0                       # softplus threshold branch
    dout = torch.randn(b, uch, L, device=dev, generator=gen)
    pr = ops.ScanProblem(u, dl, A, B, C, D, bias, softplus, hw=hw, dirs=dirs, u_mod=u_mod)
    out, x = pr.forward(True)
    grads = pr.backward(dout, x)
    grads_nockpt = pr.backward(dout, None)                  # checkpoints recomputed by an extra forward sweep

    dpg = dt // G
    flip = [k == 3 for k in dirs] if dirs is not None else [False] * G
    rep = dt // uch
    n = lambda t: None if t is None else t.float().cpu().numpy()
    un, dyn = np.tile(n(u), (1, rep, 1)), np.tile(n(dout), (1, rep, 1))
    args = (_flip_groups(un, dpg, flip), _flip_groups(n(dl), dpg, flip), n(A), _flip_groups(n(B), 1, flip),
            _flip_groups(n(C), 1, flip), n(D), n(bias))
    ref_out, ref_last = c_oracle.scan_fwd(*args, softplus, acc="f64")
    ref = c_oracle.scan_bwd(*args, _flip_groups(dyn, dpg, flip), softplus, acc="f64")
    ref_out = _flip_groups(ref_out, dpg, flip)
    ref["du"], ref["ddelta"] = _flip_groups(ref["du"], dpg, flip), _flip_groups(ref["ddelta"], dpg, flip)
    ref["dB"], ref["dC"] = _flip_groups(ref["dB"], 1, flip), _flip_groups(ref["dC"], 1, flip)

    def rel(got, want):
        got = np.asarray(n(got), np.float64)
        assert got.shape == want.shape, (got.shape, want.shape)
        return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))

    assert rel(out, ref_out) < 1e-3
    assert rel(x[:, :, -1, 1::2], ref_last) < 1e-3
    for name, g, g2 in zip(NAMES, grads, grads_nockpt):
        if ref[name] is None:
            assert g is None and g2 is None, name
            continue
        assert rel(g, ref[name]) < 1e-3, name               # du stays per direction when u is shared: same layout as ref
        assert rel(g2, ref[name]) < 1e-3, name + " (no checkpoints)"


@pytest.mark.parametrize("shape", [
    (2, 64, 256, 16, 4),        # 16 rows per group: two of the four warps of a CTA idle
    (1, 160, 100, 16, 4),       # 40 rows per group (ragged last CTA), tail of 4 positions
    (2, 32, 64, 12, 2),         # padded states (12 of 16)
    (1, 192, 512, 9, 1),        # smallest state count on the fast path; six CTAs share dB / dC (vector reductions)
], ids=str)
def test_fast_path_scan_layout(shape):
    _case(*shape)


def test_fast_decay_long_rows():
    """A of the size real checkpoints have (A = -(1 .. N)): carried-in states die within a few positions — the segmented
    mode's fix-up pass leaves a segment early once every further term is exactly zero."""
    _case(1, 64, 2048, 16, 2, a_scale=30.0)
    _case(1, 32, 1024, 12, 1, hw=(32, 32), dirs=[3], a_scale=8.0)


def test_fast_path_no_softplus():
    _case(1, 64, 128, 16, 2, softplus=False)


def test_fast_path_directions_1_3():
    _case(2, 96, 96, 16, 4, hw=(8, 12), dirs=[1, 3, 1, 3])


def test_fast_path_reversed_with_tail_and_shared_input():
    # L = 100: the last tile of the reversed traversal starts 28 positions before the row (TMA zero-fill on load)
    _case(2, 48, 100, 16, 2, hw=(10, 10), dirs=[1, 3], u_mod=24)


def test_fast_path_reversed_long():
    _case(1, 40, 1000, 16, 1, hw=(25, 40), dirs=[3])


def _fuzz_cases():
    rng = np.random.RandomState(4321)
    out = []
    for i in range(24):
        G = int(rng.choice([1, 2, 3, 4]))
        dpg = int(rng.choice([1, 2, 7, 8, 9, 16, 24, 31, 32, 33, 48, 70]))
        N = int(rng.randint(9, 17))
        H, W = int(rng.choice([2, 4, 6, 8, 10, 14, 20])), int(rng.choice([2, 4, 6, 10, 16, 18]))
        nb = int(rng.choice([1, 2, 3]))
        natural = bool(rng.randint(2))
        dirs = [int(rng.choice([1, 3])) for _ in range(G)] if natural else None
        shared = natural and bool(rng.randint(2))
        out.append(dict(b=nb, dt=G * dpg, L=H * W, N=N, G=G, hw=(H, W) if natural else None, dirs=dirs,
                        u_mod=dpg if shared else 0, softplus=bool(rng.randint(4) > 0), seed=i,
                        has_D=bool(rng.randint(4) > 0), has_bias=bool(rng.randint(4) > 0)))
    return out


@pytest.mark.parametrize("kw", _fuzz_cases(), ids=lambda k: "b%d_d%d_l%d_n%d_g%d_%s_u%d_%d" % (
    k["b"], k["dt"], k["L"], k["N"], k["G"], "".join(map(str, k["dirs"])) if k["dirs"] else "scan", k["u_mod"], k["seed"]))
def test_fast_path_fuzz(kw):
    _case(**kw)


def test_small_batch_segmented_forward_equals_sequential(forced_path):
    """Batch 1 of the north-star shape picks the segmented forward by itself (24 CTAs of sequential work otherwise): outputs,
    last state and the gradients computed from its (fixed-up) checkpoints against the sequential 8-row-warp forward."""
    if forced_path != 1:
        pytest.skip("policy set inside the test")
    from ceigm_unet_b200 import _lib, ops
    gen = torch.Generator(device="cuda").manual_seed(3)
    b, dt, L, N, G = 1, 768, 3136, 16, 4
    A = -0.5 * torch.rand(dt, N, device="cuda", generator=gen)
    B = torch.randn(b, G, N, L, device="cuda", generator=gen)
    C = torch.randn(b, G, N, L, device="cuda", generator=gen)
    D = torch.randn(dt, device="cuda", generator=gen)
    bias = 0.5 * torch.rand(dt, device="cuda", generator=gen)
    u = torch.randn(b, dt, L, device="cuda", generator=gen)
    dl = 0.5 * torch.rand(b, dt, L, device="cuda", generator=gen)
    dout = torch.randn(b, dt, L, device="cuda", generator=gen)
    res = {}
    for policy in (0, 3):
        _lib.test_force_path(policy)
        pr = ops.ScanProblem(u, dl, A, B, C, D, bias, True)
        out, x = pr.forward(True)
        res[policy] = (out, x[:, :, -1, 1::2].clone()) + tuple(pr.backward(dout, x))
    for got, want in zip(res[0], res[3]):
        err = float((got.double() - want.double()).abs().max() / want.double().abs().max().clamp_min(1e-30))
        assert err < 1e-4, err


def test_full_size_adjoint_identity():
    """Size-independent property at the BASELINE config-4 point (B=24, K=4, D=192, L=56^2, N=16): the backward is the
    adjoint of the forward's linear map u -> out, <dout, J u2> = <J^T dout, u2>."""
    from ceigm_unet_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(1)
    b, dt, L, N, G = 24, 768, 3136, 16, 4
    dev = "cuda"
    A = -0.5 * torch.rand(dt, N, device=dev, generator=gen)
    B = torch.randn(b, G, N, L, device=dev, generator=gen)
    C = torch.randn(b, G, N, L, device=dev, generator=gen)
    D = torch.randn(dt, device=dev, generator=gen)
    bias = 0.5 * torch.rand(dt, device=dev, generator=gen)
    dl = 0.5 * torch.rand(b, dt, L, device=dev, generator=gen)
    u1 = torch.randn(b, dt, L, device=dev, generator=gen)
    u2 = torch.randn(b, dt, L, device=dev, generator=gen)
    dout = torch.randn(b, dt, L, device=dev, generator=gen)
    p1 = ops.ScanProblem(u1, dl, A, B, C, D, bias, True)
    out1, x1 = p1.forward(True)
    du = p1.backward(dout, x1)[0]
    out2, _ = ops.ScanProblem(u2, dl, A, B, C, D, bias, True).forward(False)
    out0, _ = ops.ScanProblem(torch.zeros_like(u2), dl, A, B, C, D, bias, True).forward(False)
    lhs = float((dout.double() * (out2.double() - out0.double())).sum())        # <dout, J u2>
    rhs = float((du.double() * u2.double()).sum())                              # <J^T dout, u2>
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0)
