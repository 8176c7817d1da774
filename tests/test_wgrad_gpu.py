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
