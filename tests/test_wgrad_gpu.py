"""GPU parity of the tall-skinny weight-gradient kernel (csrc/wgrad.cu, C ABI ss2d_wgrad_ts) and of the autograd
Functions that use it (functional.linear_ts / proj_cm) against torch's own fp64 / autograd results.
Tolerance: rel <= 1e-3 fp32 (sums of up to 75 264 products), 2e-2 for 16-bit operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("B,R,M,N", [(24, 3136, 32, 16), (3, 777, 5, 87), (2, 100, 64, 64), (1, 33, 1, 1), (4, 196, 112, 7),
                                      (2, 4000, 3, 16)])
@pytest.mark.parametrize("layout", ["rows", "channels"])
def test_wgrad_ts_matches_einsum(B, R, M, N, layout):
    from ceigm_unet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + R + M + N)
    if layout == "rows":            # channels-last rows of a Linear: (B, R, C)
        dy = torch.randn(B, R, M, device="cuda", generator=g)
        x = torch.randn(B, R, N, device="cuda", generator=g)
    else:                           # channel-major (B, C, L) operands of the 1 x 1 projections, passed as transposed views
        dy = torch.randn(B, M, R, device="cuda", generator=g).transpose(1, 2)
        x = torch.randn(B, N, R, device="cuda", generator=g).transpose(1, 2)
    want = torch.einsum("brm,brn->mn", dy.double(), x.double())
    got = ops.wgrad_ts(dy, x)
    assert got.shape == (M, N) and got.dtype == torch.float32
    assert rel(got, want) < 1e-3
    assert torch.equal(got, ops.wgrad_ts(dy, x))               # deterministic (fixed summation order)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_wgrad_ts_16bit_operands(dtype):
    from ceigm_unet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    dy = torch.randn(2, 2048, 32, device="cuda", generator=g).to(dtype)
    x = torch.randn(2, 2048, 16, device="cuda", generator=g)            # mixed: 16-bit gradient, fp32 activation
    want = torch.einsum("brm,brn->mn", dy.double(), x.double())
    assert rel(ops.wgrad_ts(dy, x), want) < 2e-2


def test_wgrad_ts_rejects_large_tiles():
    from ceigm_unet_b200 import ops
    dy, x = torch.zeros(1, 8, 300, device="cuda"), torch.zeros(1, 8, 4, device="cuda")
    assert not ops.wgrad_ts_supported(300, 4) and not ops.wgrad_ts_supported(128, 128)
    with pytest.raises(RuntimeError):
        ops.wgrad_ts(dy, x)


def test_linear_ts_and_proj_cm_gradients_match_torch():
    from ceigm_unet_b200 import functional as Fn
    torch.manual_seed(0)
    x = torch.randn(4, 30, 30, 16, device="cuda", requires_grad=True)          # 3 600 rows: kernel path
    W = torch.randn(32, 16, device="cuda", requires_grad=True)
    b = torch.randn(32, device="cuda", requires_grad=True)
    gy = torch.randn(4, 30, 30, 32, device="cuda")
    y = Fn.linear_ts(x, W, b)
    gx, gW, gb = torch.autograd.grad(y, (x, W, b), gy)
    y0 = torch.nn.functional.linear(x, W, b)
    gx0, gW0, gb0 = torch.autograd.grad(y0, (x, W, b), gy)
    assert rel(y, y0) < 1e-6 and rel(gx, gx0) < 1e-4 and rel(gW, gW0) < 1e-3 and rel(gb, gb0) < 1e-4

    u = torch.randn(4, 16, 900, device="cuda", requires_grad=True)
    Wp = torch.randn(3, 16, device="cuda", requires_grad=True)
    go = torch.randn(4, 3, 900, device="cuda")
    o = Fn.proj_cm(Wp, u)
    gWp, gu = torch.autograd.grad(o, (Wp, u), go)
    o0 = torch.matmul(Wp, u)
    gWp0, gu0 = torch.autograd.grad(o0, (Wp, u), go)
    assert rel(o, o0) < 1e-6 and rel(gu, gu0) < 1e-4 and rel(gWp, gWp0) < 1e-3


@pytest.mark.parametrize("rows,C", [(75264, 64), (18816, 128), (4704, 348), (1176, 448), (7, 5), (33, 512), (100, 33)])
def test_layer_norm_rows_matches_torch(rows, C):
    """Row-wise LayerNorm kernel (csrc/layernorm.cu) against F.layer_norm in fp64: output and all three gradients."""
    from ceigm_unet_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(rows + C)
    x = (torch.randn(rows, C, device="cuda", generator=g) * 2 + 0.5).requires_grad_(True)
    w = torch.randn(C, device="cuda", generator=g).requires_grad_(True)
    b = torch.randn(C, device="cuda", generator=g).requires_grad_(True)
    gy = torch.randn(rows, C, device="cuda", generator=g)
    y = Fn.layer_norm_rows(x, w, b, 1e-5)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
    x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    y0 = torch.nn.functional.layer_norm(x64, (C,), w64, b64, 1e-5)
    gx0, gw0, gb0 = torch.autograd.grad(y0, (x64, w64, b64), gy.double())
    assert rel(y, y0.detach()) < 1e-5 and rel(gx, gx0) < 1e-4 and rel(gw, gw0) < 1e-3 and rel(gb, gb0) < 1e-3


def test_layer_norm_rows_bf16_and_fallback():
    from ceigm_unet_b200 import functional as Fn
    x = torch.randn(300, 96, device="cuda", dtype=torch.bfloat16)
    w, b = torch.rand(96, device="cuda") + 0.5, torch.randn(96, device="cuda")
    y = Fn.layer_norm_rows(x, w, b, 1e-5)
    y0 = torch.nn.functional.layer_norm(x.float(), (96,), w, b, 1e-5)
    assert y.dtype == torch.bfloat16 and rel(y.float(), y0) < 2e-2
    xl = torch.randn(10, 600, device="cuda")              # C > 512: library LayerNorm
    wl, bl = torch.ones(600, device="cuda"), torch.zeros(600, device="cuda")
    assert rel(Fn.layer_norm_rows(xl, wl, bl, 1e-5), torch.nn.functional.layer_norm(xl, (600,), wl, bl, 1e-5)) < 1e-6


@pytest.mark.parametrize("B,C,H,W", [(24, 16, 56, 56), (3, 87, 14, 14), (2, 5, 7, 9), (1, 1, 1, 1), (2, 192, 20, 12)])
def test_dwconv3_parameter_gradients_match_torch(B, C, H, W):
    """ss2d_dwconv3_wgrad against autograd of F.conv2d (fp64): weight and bias gradients of the depthwise 3 x 3 conv."""
    from ceigm_unet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + C + H + W)
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    dy = torch.randn(B, C, H, W, device="cuda", generator=g)
    Wt = torch.randn(C, 1, 3, 3, device="cuda", dtype=torch.float64, requires_grad=True)
    bt = torch.randn(C, device="cuda", dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(x.double(), Wt, bt, padding=1, groups=C)
    gW0, gb0 = torch.autograd.grad(y, (Wt, bt), dy.double())
    gW, gb = ops.dwconv3_wgrad(x, dy, True)
    assert gW.shape == (C, 1, 3, 3) and rel(gW, gW0) < 1e-3 and rel(gb, gb0) < 1e-3
    assert ops.dwconv3_wgrad(x, dy, False)[1] is None


def test_dwconv3_function_matches_conv_module():
    from ceigm_unet_b200 import functional as Fn
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(16, 16, 3, padding=1, groups=16, bias=True).cuda()
    x = torch.randn(8, 16, 40, 40, device="cuda", requires_grad=True)          # 12 800 pixels per channel: kernel path
    gy = torch.randn(8, 16, 40, 40, device="cuda")
    y = Fn.dwconv3(x, conv)
    gx, gW, gb = torch.autograd.grad(y, (x, conv.weight, conv.bias), gy)
    y0 = conv(x)
    gx0, gW0, gb0 = torch.autograd.grad(y0, (x, conv.weight, conv.bias), gy)
    assert rel(y, y0) < 1e-6 and rel(gx, gx0) < 1e-4 and rel(gW, gW0) < 1e-3 and rel(gb, gb0) < 1e-3
