"""GPU parity of cross-scan / cross-merge (bit-exact index work) and of the NATURAL-layout scan, where the
permutations are folded into the scan kernels' addressing instead of being materialised."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(got, ref):
    got = got.detach().double().cpu().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    ref = ref.detach().double().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def test_cross_functions_bit_exact_vs_reference_golden():
    import ceigm_unet_b200 as P
    z = np.load(os.path.join(GOLDEN, "cross_scan_merge.npz"))
    g = {k: z[k] for k in z.files}
    for tag in ("6x6", "4x6"):
        x, ys = torch.from_numpy(g[f"x_{tag}"]).cuda(), torch.from_numpy(g[f"ysK_{tag}"]).cuda()
        assert np.array_equal(P.CrossScan.apply(x).cpu().numpy(), g[f"scanK_{tag}"])
        assert np.array_equal(P.CrossMerge.apply(ys).cpu().numpy(), g[f"mergeK_{tag}"])
        xg = x.clone().requires_grad_(True)
        (P.CrossScan.apply(xg) * torch.from_numpy(g[f"scanK_bwd_w_{tag}"]).cuda()).sum().backward()
        assert np.array_equal(xg.grad.cpu().numpy(), g[f"scanK_bwd_{tag}"])
        yg = ys.clone().requires_grad_(True)
        (P.CrossMerge.apply(yg) * torch.from_numpy(g[f"mergeK_bwd_w_{tag}"]).cuda()).sum().backward()
        assert np.array_equal(yg.grad.cpu().numpy(), g[f"mergeK_bwd_{tag}"])
        for k in (1, 2, 3, 4):
            S, M = getattr(P, f"CrossScan_{k}"), getattr(P, f"CrossMerge_{k}")
            assert np.array_equal(S.apply(x).cpu().numpy(), g[f"scan{k}_{tag}"])
            assert np.array_equal(M.apply(ys[:, k - 1:k].contiguous()).cpu().numpy(), g[f"merge{k}_{tag}"])
            if tag == "6x6":        # the reference's _2/_4 backward is a true adjoint only on square maps
                xg = x.clone().requires_grad_(True)
                (S.apply(xg) * torch.from_numpy(g[f"scan{k}_bwd_w_{tag}"]).cuda()).sum().backward()
                assert np.array_equal(xg.grad.cpu().numpy(), g[f"scan{k}_bwd_{tag}"])
                yg = ys[:, k - 1:k].clone().requires_grad_(True)
                (M.apply(yg) * torch.from_numpy(g[f"merge{k}_bwd_w_{tag}"]).cuda()).sum().backward()
                assert np.array_equal(yg.grad.cpu().numpy(), g[f"merge{k}_bwd_{tag}"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("hw", [(56, 56), (14, 14), (7, 7), (56, 57), (5, 9)])
def test_cross_permutations_vs_oracle(hw, dtype):
    """Integer-valued inputs: every output element must equal the oracle's exactly (CHECK.check_csm_triton uses 56x57)."""
    import ceigm_unet_b200 as P
    from oracle import cross_scan as O
    H, W = hw
    x = (torch.arange(2 * 5 * H * W) % 251).to(dtype).view(2, 5, H, W)
    ys = (torch.arange(2 * 4 * 5 * H * W) % 127).to(dtype).view(2, 4, 5, H, W)
    assert torch.equal(P.CrossScan.apply(x.cuda()).cpu(), O.cross_scan4(x))
    assert torch.equal(P.CrossMerge.apply(ys.cuda()).cpu(), O.cross_merge4(ys))
    for k in (1, 2, 3, 4):
        assert torch.equal(getattr(P, f"CrossScan_{k}").apply(x.cuda()).cpu(), O.cross_scan_k(x, k))
        assert torch.equal(getattr(P, f"CrossMerge_{k}").apply(ys[:, k - 1:k].contiguous().cuda()).cpu(),
                           O.cross_merge_k(ys[:, k - 1:k], k))
        # merge_k(scan_k(x)) == x bit-exactly
        back = getattr(P, f"CrossMerge_{k}").apply(getattr(P, f"CrossScan_{k}").apply(x.cuda()).view(2, 1, 5, H, W))
        assert torch.equal(back.cpu().view_as(x), x)


@pytest.mark.parametrize("hw,dirs,N,D", [((56, 56), (1, 2, 3, 4), 16, 24), ((14, 14), (1, 2, 3, 4), 16, 20),
                                         ((7, 7), (4,), 1, 28), ((14, 14), (2,), 1, 87), ((9, 12), (1, 2, 3, 4), 4, 8),
                                         ((28, 28), (3,), 1, 32), ((28, 28), (1,), 1, 32)])
def test_natural_layout_scan_equals_permute_scan_merge(hw, dirs, N, D):
    """NATURAL layout + directions == cross_scan -> scan (SCAN layout) -> per-direction un-permute, forward and backward,
    against the f64 oracle composition."""
    from ceigm_unet_b200 import ops
    from oracle import c_oracle
    from oracle import cross_scan as O
    H, W = hw
    K, L, Bn = len(dirs), H * W, 2
    gen = torch.Generator().manual_seed(H * 100 + W + N)
    x = torch.randn(Bn, D, H, W, generator=gen)
    dts = 0.5 * torch.rand(Bn, K * D, H, W, generator=gen)
    Bs, Cs = torch.randn(Bn, K, N, H, W, generator=gen), torch.randn(Bn, K, N, H, W, generator=gen)
    A = -0.5 * torch.rand(K * D, N, generator=gen)
    Dv, bias = torch.randn(K * D, generator=gen), 0.5 * torch.rand(K * D, generator=gen)
    dy = torch.randn(Bn, D, H, W, generator=gen)           # merged gradient, shared by all directions

    # oracle: materialise the permutations on the CPU
    def perm(t, k):     # (B, C, H, W) natural -> (B, C, L) in direction k
        return O.cross_scan_k(t, k)[:, 0]
    xs = torch.cat([perm(x, k) for k in dirs], 1)
    dls = torch.cat([perm(dts[:, i * D:(i + 1) * D], k) for i, k in enumerate(dirs)], 1)
    Bp = torch.stack([perm(Bs[:, i], k) for i, k in enumerate(dirs)], 1)
    Cp = torch.stack([perm(Cs[:, i], k) for i, k in enumerate(dirs)], 1)
    dys = torch.cat([perm(dy, k) for k in dirs], 1)
    n = lambda t: t.numpy()
    ref_out, _ = c_oracle.scan_fwd(n(xs), n(dls), n(A), n(Bp), n(Cp), n(Dv), n(bias), True)
    ref_g = c_oracle.scan_bwd(n(xs), n(dls), n(A), n(Bp), n(Cp), n(Dv), n(bias), n(dys), True)

    def unperm(a, k):   # (B, C, L) scan order -> (B, C, L) natural order
        t = torch.from_numpy(np.ascontiguousarray(a))
        return O.cross_merge_k(t.view(t.shape[0], 1, t.shape[1], H, W), k)

    prob = ops.ScanProblem(x.view(Bn, D, L).cuda(), dts.view(Bn, K * D, L).cuda(), A.cuda(), Bs.view(Bn, K, N, L).cuda(),
                           Cs.view(Bn, K, N, L).cuda(), Dv.cuda(), bias.cuda(), True, out_float=True, hw=hw, dirs=dirs, u_mod=D)
    out, st = prob.forward(True)
    for i, k in enumerate(dirs):
        assert rel_err(out[:, i * D:(i + 1) * D], unperm(ref_out[:, i * D:(i + 1) * D], k)) < 1e-3
    du, ddl, dA, dB, dC, dD, dbias = prob.backward(dy.view(Bn, D, L).cuda(), st)
    for i, k in enumerate(dirs):
        sl = slice(i * D, (i + 1) * D)
        assert rel_err(du[:, sl], unperm(ref_g["du"][:, sl], k)) < 1e-3
        assert rel_err(ddl[:, sl], unperm(ref_g["ddelta"][:, sl], k)) < 1e-3
        assert rel_err(dB[:, i], unperm(ref_g["dB"][:, i], k)) < 1e-3
        assert rel_err(dC[:, i], unperm(ref_g["dC"][:, i], k)) < 1e-3
    assert rel_err(dA, ref_g["dA"]) < 1e-3
    assert rel_err(dD, ref_g["dD"]) < 1e-3
    assert rel_err(dbias, ref_g["ddelta_bias"]) < 1e-3
