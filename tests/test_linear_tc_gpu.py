"""GPU parity tests of the tensor-core projection kernel (csrc/linear_tc.cu, C ABI ss2d_linear_tc) that replaces
in_proj + chunk + the NHWC -> NCHW copy (model/gm/ss2d.py:504-510) and out_proj (ss2d.py:518).

Reference: the same product in fp64 (torch.matmul on the CPU-exact operands). Tolerances: fp32 operands run as TF32 on
the tensor cores;
bf16 operands are exact in the multiplier, so only the bf16 rounding of the output remains: rel <= 2e-2 (bar), measured
against the fp64 product of the SAME bf16 operands it is ~4e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,K,N", [(128, 32, 16), (1000, 96, 256), (24 * 196, 64, 128), (300, 192, 96), (5, 160, 256)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rows_single_part(M, K, N, dtype):
    from ceigm_unet_b200 import ops
    if not ops.linear_tc_supported(N, K, dtype):
        pytest.skip("part does not fit")
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(dtype)
    b = torch.randn(N, device="cuda", generator=g)
    (out,) = ops.linear_tc(x, W, b, [(N, "rows", False)])
    ref = x.double() @ W.double().t() + b.double()
    assert out.shape == (M, N) and out.dtype == dtype
    assert _rel(out, ref) < (4e-4 if dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_in_proj_split_planes_and_gate(dtype):
    """in_proj of SS2D(d_model=96): x half -> (B, D, H, W) planes, z half -> rows; batch boundary inside a 128-row tile."""
    from ceigm_unet_b200 import ops
    Bn, H, Wd, C, D = 3, 14, 10, 96, 192
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(Bn, H, Wd, C, device="cuda", generator=g).to(dtype)
    W = (torch.randn(2 * D, C, device="cuda", generator=g) / C ** 0.5).to(dtype)
    xi, z = ops.linear_tc(x, W, None, [(D, ("planes", H * Wd), False), (D, "rows", True)])
    ref = x.double().reshape(-1, C) @ W.double().t()
    ref_x = ref[:, :D].reshape(Bn, H * Wd, D).transpose(1, 2)
    ref_z = torch.nn.functional.silu(ref[:, D:]).reshape(Bn, H, Wd, D)
    assert xi.shape == (Bn, D, H * Wd) and z.shape == (Bn, H, Wd, D)
    tol = 4e-4 if dtype == torch.float32 else 6e-3
    assert _rel(xi, ref_x) < tol and _rel(z, ref_z) < tol


def test_strided_rows_and_many_tiles():
    """A is a column slice of a wider matrix (row stride != K); more tiles than CTAs."""
    from ceigm_unet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    big = torch.randn(40000, 128, device="cuda", generator=g)
    x = big[:, 32:96]
    W = torch.randn(64, 64, device="cuda", generator=g) / 8
    (out,) = ops.linear_tc(x, W, None, [(64, "rows", False)])
    assert _rel(out, x.double() @ W.double().t()) < 4e-4


def test_rejects_bad_shapes():
    from ceigm_unet_b200 import ops
    x = torch.randn(64, 32, device="cuda")
    W = torch.randn(24, 32, device="cuda")
    with pytest.raises(RuntimeError):
        ops.linear_tc(x, W, None, [(24, "rows", False)])          # n_cols not a multiple of 16
    with pytest.raises(RuntimeError):
        ops.linear_tc(x.cpu(), W.cpu(), None, [(24, "rows", False)])


def test_no_write_outside_the_outputs():
    """Guard bands around both output parts (rows and planes), M not a multiple of the 128-row tile: the epilogue's row
    clipping must keep every store inside [0, M) — checked through the C ABI with outputs placed inside sentinel buffers."""
    import ctypes
    from ceigm_unet_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    Bn, Lp, K, D = 3, 50, 64, 48                       # M = 150: one full tile + a 22-row tail
    M = Bn * Lp
    x = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(2 * D, K, device="cuda", generator=g) / 8
    PAD = 4096
    buf_p = torch.full((PAD + Bn * D * Lp + PAD,), 777.0, device="cuda")
    buf_r = torch.full((PAD + M * D + PAD,), 777.0, device="cuda")
    parts = (_lib.LinearPart * 2)()
    parts[0].out, parts[0].ld, parts[0].n_cols, parts[0].planes_L, parts[0].act = buf_p.data_ptr() + 4 * PAD, 0, D, Lp, 0
    parts[1].out, parts[1].ld, parts[1].n_cols, parts[1].planes_L, parts[1].act = buf_r.data_ptr() + 4 * PAD, D, D, 0, 0
    rc = L.ss2d_linear_tc(ctypes.c_void_p(x.data_ptr()), K, ctypes.c_void_p(W.data_ptr()), K, None, M, 2 * D, K, _lib.SS2D_F32, 2, parts,
                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    ref = x.double() @ W.double().t()
    for buf, n in ((buf_p, Bn * D * Lp), (buf_r, M * D)):
        assert bool((buf[:PAD] == 777.0).all()) and bool((buf[PAD + n:] == 777.0).all())
    planes = buf_p[PAD:PAD + Bn * D * Lp].view(Bn, D, Lp)
    rows = buf_r[PAD:PAD + M * D].view(M, D)
    assert _rel(planes, ref[:, :D].reshape(Bn, Lp, D).transpose(1, 2)) < 4e-4
    assert _rel(rows, ref[:, D:]) < 4e-4
