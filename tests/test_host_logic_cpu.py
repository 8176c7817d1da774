"""CPU tests of the host-side helpers around the kernels (no GPU, no CUDA extension calls): the strided-batched and split-K
GEMM formulations used by modules.SS2D (functional._bmm_w / _wgrad_rows) against plain matmul, the block-structured x_proj weight
of the two-plane K = 4 path against the per-direction projections of model/gm/ss2d.py:465-466, and the eligibility predicates
that must refuse CPU tensors instead of falling back."""
import pytest
import torch
import torch.nn.functional as F


def test_bmm_w_equals_matmul():
    from ceigm_unet_b200 import functional as Fn
    g = torch.Generator().manual_seed(0)
    W = torch.randn(38, 24, generator=g)
    u = torch.randn(3, 24, 50, generator=g)
    assert torch.allclose(Fn._bmm_w(W, u), torch.matmul(W, u), atol=1e-5)
    assert torch.allclose(Fn._bmm_w(W.t(), torch.matmul(W, u)), torch.matmul(W.t(), torch.matmul(W, u)), atol=1e-4)


@pytest.mark.parametrize("rows", [256 * 128, 1000, 3 * 4096])
def test_wgrad_rows_equals_one_gemm(rows):
    from ceigm_unet_b200 import functional as Fn
    g = torch.Generator().manual_seed(rows)
    dy = torch.randn(rows, 12, generator=g, dtype=torch.float64)
    x = torch.randn(rows, 7, generator=g, dtype=torch.float64)
    assert torch.allclose(Fn._wgrad_rows(dy, x), dy.t() @ x, rtol=1e-10, atol=1e-9)


def test_block_structured_x_proj_equals_per_direction_projections():
    """W_blk (K C, 2 D) applied to u = [natural plane | transposed plane] gives x_dbl in group order: direction k reads
    plane k % 2 (modules.SS2D.forward_core), i.e. einsum('bkdl,kcd->bkcl') on the planes each direction scans."""
    g = torch.Generator().manual_seed(1)
    K, C, D, Bn, L = 4, 10, 6, 2, 12
    Wx = torch.randn(K, C, D, generator=g)
    planes = [torch.randn(Bn, D, L, generator=g) for _ in range(2)]
    u = torch.cat(planes, dim=1)
    W_blk = torch.cat([F.pad(Wx[k], ((k % 2) * D, (1 - k % 2) * D)) for k in range(K)])
    got = torch.matmul(W_blk, u).view(Bn, K, C, L)
    xs = torch.stack([planes[k % 2] for k in range(K)], dim=1)
    ref = torch.einsum("bkdl,kcd->bkcl", xs, Wx)
    assert torch.allclose(got, ref, atol=1e-5)


def test_tensor_core_paths_refuse_cpu_tensors():
    from ceigm_unet_b200 import functional as Fn
    x = torch.randn(1, 4, 4, 16)
    W = torch.randn(32, 16)
    with pytest.raises(RuntimeError):
        Fn.in_proj_planes(x, W)
    with pytest.raises(RuntimeError):
        Fn.linear_tc(x, W)
    with pytest.raises(RuntimeError):
        Fn.dwconv3_silu_planes(torch.randn(1, 4, 8, 8), torch.randn(4, 1, 3, 3), None)


def test_harness_min_pool_rebinding_matches_reference_on_cpu():
    """harness.graph_step._min_pool_direct vs the reference's AdaptiveMinPool2d.forward (F.unfold of the whole map + min,
    model/best_decoder.py:179-191): same values, same gradient routing (ties included). CPU tensors: host logic only."""
    import importlib
    import pytest
    from harness import graph_step, refmodel
    if not refmodel.available():
        pytest.skip("reference tree not installed (harness/install_ref.py)")
    refmodel.load_reference(scan="cpu_fast")
    bd = importlib.import_module("model.best_decoder")
    orig = getattr(bd.AdaptiveMinPool2d, "_ss2d_harness_orig_forward", bd.AdaptiveMinPool2d.forward)
    pool = bd.AdaptiveMinPool2d()
    gen = torch.Generator().manual_seed(0)
    x = torch.relu(torch.randn(2, 6, 7, 7, generator=gen))
    x[:, ::2] -= 1.0
    g = torch.randn(2, 6, 1, 1, generator=gen)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    orig(pool, xa).backward(g)
    graph_step._min_pool_direct(pool, xb).backward(g)
    assert torch.equal(orig(pool, x), graph_step._min_pool_direct(pool, x)) and torch.equal(xa.grad, xb.grad)
