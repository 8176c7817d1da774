"""GPU parity of the fused depthwise-3x3 + SiLU kernels (csrc/dwconv.cu: ss2d_dwconv3_act; model/gm/ss2d.py:512-513)
against the library composition F.silu(F.conv2d(x, W, b, padding=1, groups=C)) and its autograd gradients.
Tolerance: rel <= 1e-3 fp32 (cuDNN TF32 convolutions disabled for the comparison), <= 2e-2 bf16."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    return float((got.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))


@pytest.fixture(autouse=True)
def _fp32_conv():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("shape", [(2, 192, 56, 56), (3, 64, 28, 28), (2, 348, 14, 14), (2, 112, 7, 7), (1, 5, 9, 13), (1, 3, 1, 4)], ids=str)
@pytest.mark.parametrize("dtype,bias", [(torch.float32, True), (torch.float32, False), (torch.bfloat16, True)])
def test_forward_and_gradients(shape, dtype, bias):
    from ceigm_unet_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    Bn, C, H, W = shape
    x = torch.randn(shape, device="cuda", generator=g).to(dtype).requires_grad_(True)
    Wt = (0.3 * torch.randn(C, 1, 3, 3, device="cuda", generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True) if bias else None
    dy = torch.randn(shape, device="cuda", generator=g).to(dtype)
    y = Fn.dwconv3_silu(x, Wt, b)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), Wt.grad.clone(), None if b is None else b.grad.clone())
    x.grad = None; Wt.grad = None
    if b is not None:
        b.grad = None
    yr = F.silu(F.conv2d(x.float(), Wt, b, padding=1, groups=C))
    yr.backward(dy.float())
    ref = (yr.detach(), x.grad, Wt.grad, None if b is None else b.grad)
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    assert y.dtype == dtype and y.shape == x.shape
    assert _rel(got[0], ref[0]) < tol and _rel(got[1], ref[1]) < tol
    assert _rel(got[2], ref[2]) < 2 * tol
    if b is not None:
        assert _rel(got[3], ref[3]) < 2 * tol


def test_no_cpu_path():
    from ceigm_unet_b200 import functional as Fn
    with pytest.raises(RuntimeError):
        Fn.dwconv3_silu(torch.randn(1, 4, 8, 8), torch.randn(4, 1, 3, 3), None)


@pytest.mark.parametrize("shape", [(2, 192, 56, 56), (2, 24, 28, 36), (1, 7, 4, 8), (3, 5, 40, 32)], ids=str)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_two_plane_output_and_its_adjoint(shape, dtype):
    """u = [SiLU(conv(x)) | the same image transposed] written by one kernel (the scan's two input planes, csms6s.py:11-29),
    and the backward that adds the two planes' gradients on the fly, against conv + SiLU + transpose + cat under autograd."""
    from ceigm_unet_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    Bn, C, H, W = shape
    x = torch.randn(shape, device="cuda", generator=g).to(dtype).requires_grad_(True)
    Wt = (0.3 * torch.randn(C, 1, 3, 3, device="cuda", generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    du = torch.randn(Bn, 2 * C, H * W, device="cuda", generator=g).to(dtype)
    u = Fn.dwconv3_silu_planes(x, Wt, b)
    assert u is not None and u.shape == (Bn, 2 * C, H * W) and u.dtype == dtype
    u.backward(du)
    got = (u.detach(), x.grad.clone(), Wt.grad.clone(), b.grad.clone())
    x.grad = None; Wt.grad = None; b.grad = None
    yr = F.silu(F.conv2d(x.float(), Wt, b, padding=1, groups=C))
    ur = torch.cat([yr.reshape(Bn, C, H * W), yr.transpose(2, 3).reshape(Bn, C, H * W)], dim=1)
    ur.backward(du.float())
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    assert _rel(got[0], ur.detach()) < tol and _rel(got[1], x.grad) < tol
    assert _rel(got[2], Wt.grad) < 2 * tol and _rel(got[3], b.grad) < 2 * tol


def test_two_plane_output_needs_multiples_of_four():
    from ceigm_unet_b200 import functional as Fn
    assert Fn.dwconv3_silu_planes(torch.randn(1, 4, 7, 7, device="cuda"), torch.randn(4, 1, 3, 3, device="cuda"), None) is None
