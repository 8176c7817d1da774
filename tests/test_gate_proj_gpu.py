"""GPU parity of the fused epilogue + output projection kernel (csrc/gate_proj_tc.cu, C ABI ss2d_gate_proj_fwd): merge over the
K direction planes (planes of directions 2 / 4 in transposed pixel order) + LayerNorm(D) + SiLU(z) gate + out_proj
(model/gm/ss2d.py:486-498, 506-508, 515-518) against the same composition in fp64 torch.
Tolerance: fp32 tensors (TF32 tensor-core math, operands rounded to nearest) rel <= 1e-3 of max|ref|; bf16 <= 2e-2."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    return float((got.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))


def _reference(ys, tmask, lnw, lnb, eps, z, W, bias, H, Wd):
    Bn, K, D, L = ys.shape
    planes = []
    for k in range(K):
        pk = ys[:, k].double()
        if (tmask >> k) & 1:
            pk = pk.view(Bn, D, Wd, H).transpose(2, 3).reshape(Bn, D, L)
        planes.append(pk)
    y = (planes[0] + planes[2]) + (planes[1] + planes[3]) if K == 4 else sum(planes)
    yn = F.layer_norm(y.transpose(1, 2), (D,), lnw.double(), lnb.double(), eps)
    g = yn * F.silu(z.double()) if z is not None else yn
    out = g @ W.double().t()
    if bias is not None:
        out = out + bias.double()
    return out, g


@pytest.mark.parametrize("shape", [(2, 4, 192, 56, 56, 96, 0b1010), (1, 4, 64, 8, 16, 32, 0b1010), (3, 4, 128, 12, 20, 64, 0b1010),
                                   (2, 2, 64, 12, 20, 16, 0b00), (1, 1, 64, 4, 12, 48, 0b1), (1, 3, 64, 20, 8, 16, 0b101), (2, 4, 192, 28, 28, 96, 0b0000)], ids=str)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gate_proj_matches_composition(shape, dtype):
    from ceigm_unet_b200 import ops
    Bn, K, D, H, Wd, C, tmask = shape
    if not ops.gate_proj_supported(D, C, K, dtype):
        pytest.skip("shape does not fit the fused kernel")
    L = H * Wd
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    ys = torch.randn(Bn, K, D, L, device="cuda", generator=g)
    lnw = 1.0 + 0.2 * torch.randn(D, device="cuda", generator=g)
    lnb = 0.1 * torch.randn(D, device="cuda", generator=g)
    z = torch.randn(Bn, L, D, device="cuda", generator=g).to(dtype)
    W = (torch.randn(C, D, device="cuda", generator=g) / D ** 0.5).to(dtype)
    bias = 0.1 * torch.randn(C, device="cuda", generator=g)
    out, stats, gt = ops.gate_proj_fwd(ys, lnw, lnb, z, True, 1e-5, W, bias, (H, Wd), tmask, True)
    torch.cuda.synchronize()
    ref_out, ref_g = _reference(ys, tmask, lnw, lnb, 1e-5, z, W, bias, H, Wd)
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    assert out.shape == (Bn, L, C) and out.dtype == dtype and gt.shape == (Bn, L, D)
    assert _rel(gt, ref_g) < (1e-5 if dtype == torch.float32 else 1e-2)
    assert _rel(out, ref_out) < tol
    # the statistics are what ss2d_out_gate_bwd consumes
    planes = [ys[:, k].view(Bn, D, Wd, H).transpose(2, 3).reshape(Bn, D, L) if (tmask >> k) & 1 else ys[:, k] for k in range(K)]
    y = sum(p.double() for p in planes)
    assert _rel(stats[..., 0], y.mean(dim=1)) < 1e-4
    assert _rel(stats[..., 1], (y.var(dim=1, unbiased=False) + 1e-5).rsqrt()) < 1e-4


def test_gate_proj_z_view_and_no_g():
    """z as the strided second half of an in_proj output row (row stride 2 D), no gated-tensor output."""
    from ceigm_unet_b200 import ops
    Bn, K, D, H, Wd, C = 2, 4, 64, 8, 12, 32
    L = H * Wd
    g = torch.Generator(device="cuda").manual_seed(3)
    ys = torch.randn(Bn, K, D, L, device="cuda", generator=g)
    xz = torch.randn(Bn, L, 2 * D, device="cuda", generator=g)
    z = xz[..., D:]
    lnw, lnb = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    W = torch.randn(C, D, device="cuda", generator=g) / 8
    out, stats, gt = ops.gate_proj_fwd(ys, lnw, lnb, z, True, 1e-5, W, None, (H, Wd), 0b1010, False)
    assert gt is None
    ref_out, _ = _reference(ys, 0b1010, lnw, lnb, 1e-5, z, W, None, H, Wd)
    assert _rel(out, ref_out) < 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 4, 192, 56, 56, 0b1010), (1, 2, 64, 8, 20, 0b01), (2, 1, 128, 12, 12, 0b0)], ids=str)
def test_epilogue_only_mode_serves_out_gate_fwd(shape, dtype):
    """ops.out_gate_fwd routes eligible shapes (H, W multiples of 4, D a multiple of 64, <= 256) to the TMA-fed kernel with
    C = 0 (no projection): the gated tensor and the LayerNorm statistics against the fp64 composition."""
    from ceigm_unet_b200 import ops
    Bn, K, D, H, Wd, tmask = shape
    L = H * Wd
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    ys = torch.randn(Bn, K, D, L, device="cuda", generator=g)
    lnw = 1.0 + 0.2 * torch.randn(D, device="cuda", generator=g)
    lnb = 0.1 * torch.randn(D, device="cuda", generator=g)
    z = torch.randn(Bn, L, D, device="cuda", generator=g).to(dtype)
    assert ops.gate_proj_supported(D, 0, K, dtype)
    out, stats = ops.out_gate_fwd(ys, lnw, lnb, z, True, 1e-5, dtype, (H, Wd), tmask)
    planes = [ys[:, k].double().view(Bn, D, Wd, H).transpose(2, 3).reshape(Bn, D, L) if (tmask >> k) & 1 else ys[:, k].double() for k in range(K)]
    y = sum(planes)
    ref = F.layer_norm(y.transpose(1, 2), (D,), lnw.double(), lnb.double(), 1e-5) * F.silu(z.double())
    assert out.shape == (Bn, L, D) and out.dtype == dtype
    assert _rel(out, ref) < (1e-5 if dtype == torch.float32 else 1e-2)
    assert _rel(stats[..., 0], y.mean(dim=1)) < 1e-4
