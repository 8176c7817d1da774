"""GPU parity of the nn.Module surface against outputs and gradients recorded from the UNMODIFIED reference modules
(tests/golden/ss2d_*.npz, group_mamba_layer.npz): weights are loaded through the reference's own state_dict keys."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(got, ref):
    got = got.detach().double().cpu().numpy()
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def _check(mod, g, run, tol=1e-3):
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}
    missing, unexpected = mod.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    mod = mod.cuda()
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    y = run(mod, x)
    assert rel_err(y, g["y"]) < tol
    (y * torch.from_numpy(g["dy"]).cuda()).sum().backward()
    assert rel_err(x.grad, g["dx"]) < tol
    for n, p in mod.named_parameters():
        assert rel_err(p.grad, g["grad." + n]) < 2 * tol, n


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_ss2d_gm_single_direction(k):
    import ceigm_unet_b200 as P
    g = _load(f"ss2d_gm_dir{k}.npz")
    m = P.SS2D(d_model=8, d_state=1, ssm_ratio=1, d_conv=3)
    _check(m, g, lambda mod, x: mod(x, CrossScan=getattr(P, f"CrossScan_{k}"), CrossMerge=getattr(P, f"CrossMerge_{k}")))


def test_ss2d_vmamba_k4_n16_nonsquare():
    import ceigm_unet_b200 as P
    g = _load("ss2d_vm_k4_n16.npz")
    m = P.SS2D(d_model=8, d_state=16, ssm_ratio=2.0, k_group=4)
    _check(m, g, lambda mod, x: mod(x))


def test_group_mamba_layer():
    import ceigm_unet_b200 as P
    g = _load("group_mamba_layer.npz")
    m = P.GroupMambaLayer(32, 32)
    _check(m, g, lambda mod, x: mod(x, 6, 6))


def test_state_dict_keys_and_shapes_match_reference():
    import ceigm_unet_b200 as P
    g = _load("ss2d_vm_k4_n16.npz")
    m = P.SS2D(d_model=8, d_state=16, ssm_ratio=2.0, k_group=4)
    ref = {k[3:]: v.shape for k, v in g.items() if k.startswith("sd.")}
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert list(ours.keys()) == list(ref.keys())            # same names in the same order
    assert ours == {k: tuple(v) for k, v in ref.items()}
    g = _load("group_mamba_layer.npz")
    m = P.GroupMambaLayer(32, 32)
    ref = {k[3:]: tuple(v.shape) for k, v in g.items() if k.startswith("sd.")}
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == ref


def test_ss2d_bf16_autocast_vs_fp32():
    """Under bf16 autocast the scan still runs in fp32 (force_fp32, ss2d.py:287,479-480): result within 2e-2 of fp32."""
    import ceigm_unet_b200 as P
    torch.manual_seed(0)
    m = P.SS2D(d_model=32, d_state=16, ssm_ratio=2.0, k_group=4).cuda()
    x = torch.randn(2, 14, 14, 32, device="cuda")
    y32 = m(x)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y16 = m(x)
    assert y16.dtype == torch.bfloat16
    assert rel_err(y16.float(), y32.detach().cpu().numpy()) < 2e-2


def test_dropin_modules_serve_the_reference_functions():
    """install_dropin() registers our modules under the reference extension names; the reference-style autograd
    wrapper (csms6s.py:347-365, restated here) then runs on them unchanged and matches the golden vectors."""
    import sys

    import ceigm_unet_b200 as P
    P.install_dropin()
    import selective_scan_cuda_core                      # noqa: F401  (resolved from sys.modules)
    assert sys.modules["selective_scan_cuda_core"].__name__.endswith("selective_scan_cuda_core")

    class RefStyleCore(torch.autograd.Function):        # body of the reference's SelectiveScanCore
        @staticmethod
        def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False):
            ctx.delta_softplus = delta_softplus
            out, x, *rest = selective_scan_cuda_core.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, 1)
            ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
            return out

        @staticmethod
        def backward(ctx, dout):
            u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
            if dout.stride(-1) != 1:
                dout = dout.contiguous()
            grads = selective_scan_cuda_core.bwd(u, delta, A, B, C, D, delta_bias, dout, x, ctx.delta_softplus, 1)
            return (*grads[:7], None)

    g = _load("scan_n16_g4.npz")
    t = {k[3:]: torch.from_numpy(v).cuda().requires_grad_(k != "in_dout") for k, v in g.items() if k.startswith("in_")}
    out = RefStyleCore.apply(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["delta_bias"], True)
    out.backward(t["dout"])
    assert rel_err(out, g["out"]) < 1e-3
    for k in ("u", "delta", "A", "B", "C", "D", "delta_bias"):
        assert rel_err(t[k].grad, g["grad_" + k]) < 1e-3, k


def test_group_mamba_layer_cuda_graphs_match_eager():
    """P.graphed(): forward and backward of a whole GroupMambaLayer replayed from CUDA graphs give the eager result
    (outputs, input gradient and every parameter gradient)."""
    import copy

    import ceigm_unet_b200 as P

    class Wrap(torch.nn.Module):
        def __init__(self, layer):
            super().__init__()
            self.layer = layer

        def forward(self, x):
            return self.layer(x, 14, 14)

    torch.manual_seed(0)
    eager = Wrap(P.GroupMambaLayer(64, 64)).cuda()
    captured = copy.deepcopy(eager)               # graphed() must see the module before its first eager backward
    x = torch.randn(3, 196, 64, device="cuda", requires_grad=True)
    gy = torch.randn(3, 196, 64, device="cuda")
    g = P.graphed(captured, (x.detach().clone().requires_grad_(True),))
    outs = []
    for fn, mod in ((g, captured), (eager, eager)):
        for _ in range(2):                        # replay twice: static buffers are reused
            x.grad = None
            for p_ in mod.parameters():
                p_.grad = None
            y = fn(x)
            y.backward(gy)
        outs.append((y.detach().clone(), x.grad.clone(), [p_.grad.clone() for p_ in mod.parameters()]))
    (y1, gx1, gp1), (y2, gx2, gp2) = outs
    assert rel_err(y1, y2.cpu().numpy()) < 1e-5
    assert rel_err(gx1, gx2.cpu().numpy()) < 1e-5
    for a, b in zip(gp1, gp2):
        assert rel_err(a, b.cpu().numpy()) < 1e-4


@pytest.mark.parametrize("D", [16, 40])          # 16: thread-per-pixel epilogue kernels; 40: tiled kernels (K = 1, transposed plane)
@pytest.mark.parametrize("pair", [(2, 1), (4, 3)])
def test_column_major_directions_equal_row_major_on_the_transposed_image(D, pair):
    """SS2D.forward_core in direction 2 (4) on an image == direction 1 (3) on the transposed image, transposed back —
    CrossScan_2/_4 are CrossScan_1/_3 of the transposed input (csms6s.py:95-129, 172-206). Checks values and the
    gradients w.r.t. the input and the gate on a non-square map, through both epilogue kernel families."""
    import ceigm_unet_b200 as P
    torch.manual_seed(D)
    col, row = pair
    m = P.SS2D(d_model=D, d_state=1, ssm_ratio=1, d_conv=3).cuda()
    Bn, H, W = 2, 6, 10
    x = torch.randn(Bn, D, H, W, device="cuda", requires_grad=True)
    z = torch.randn(Bn, H, W, D, device="cuda", requires_grad=True)
    gy = torch.randn(Bn, H, W, D, device="cuda")
    y_c = m.forward_core(x, z, (col,))
    gx_c, gz_c = torch.autograd.grad(y_c, (x, z), gy)
    xt = x.detach().transpose(2, 3).contiguous().requires_grad_(True)
    zt = z.detach().transpose(1, 2).contiguous().requires_grad_(True)
    y_r = m.forward_core(xt, zt, (row,))                                   # (B, W, H, D)
    gx_r, gz_r = torch.autograd.grad(y_r, (xt, zt), gy.transpose(1, 2).contiguous())
    assert rel_err(y_c, y_r.transpose(1, 2).detach().cpu().numpy()) < 1e-5
    assert rel_err(gx_c, gx_r.transpose(2, 3).cpu().numpy()) < 1e-4
    assert rel_err(gz_c, gz_r.transpose(1, 2).cpu().numpy()) < 1e-4


def test_group_mamba_layer_bf16_autocast_at_live_size():
    """bf16 autocast through the module-level kernels at a size where they are all taken (tall-skinny weight gradients,
    depthwise-conv parameter gradients, LayerNorm, small-D epilogue): outputs and parameter gradients within 2e-2 /
    5e-2 of the fp32 run (bf16 GEMM and conv inputs, fp32 scan: ss2d.py:287, 479-480)."""
    import copy

    import ceigm_unet_b200 as P
    torch.manual_seed(0)
    m32 = P.GroupMambaLayer(64, 64).cuda()
    m16 = copy.deepcopy(m32)
    x = torch.randn(4, 48 * 48, 64, device="cuda")             # 9 216 rows
    gy = torch.randn(4, 48 * 48, 64, device="cuda")
    x32 = x.clone().requires_grad_(True)
    y32 = m32(x32, 48, 48)
    y32.backward(gy)
    x16 = x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y16 = m16(x16, 48, 48)
    y16.float().backward(gy)
    assert rel_err(y16.float(), y32.detach().cpu().numpy()) < 2e-2
    assert rel_err(x16.grad, x32.grad.cpu().numpy()) < 5e-2
    for (n, a), (_, b) in zip(m16.named_parameters(), m32.named_parameters()):
        assert a.grad is not None and a.grad.dtype == a.dtype and torch.isfinite(a.grad).all(), n
        assert rel_err(a.grad, b.grad.cpu().numpy()) < 8e-2, n


def test_ss2d_vmamba_regime_cuda_graphs_match_eager():
    """graphed() around the K = 4, d_state = 16 SS2D: the TMA tensor maps of scan_fwd / scan_bwd2 are kernel parameters
    baked at capture time over the graph's static buffers; replays must reproduce the eager forward and backward."""
    import copy

    import ceigm_unet_b200 as P
    torch.manual_seed(1)
    eager = P.SS2D(d_model=24, d_state=16, ssm_ratio=2.0, k_group=4).cuda()
    captured = copy.deepcopy(eager)
    x = torch.randn(3, 12, 16, 24, device="cuda", requires_grad=True)
    gy = torch.randn(3, 12, 16, 24, device="cuda")
    g = P.graphed(captured, (x.detach().clone().requires_grad_(True),))
    res = []
    for fn, mod in ((g, captured), (eager, eager)):
        for rep in range(2):
            xin = (x.detach() * (1.0 + 0.5 * rep)).requires_grad_(True)         # new values, same buffers on replay
            for p_ in mod.parameters():
                p_.grad = None
            y = fn(xin)
            y.backward(gy)
        res.append((y.detach().clone(), xin.grad.clone(), [p_.grad.clone() for p_ in mod.parameters()]))
    (y1, gx1, gp1), (y2, gx2, gp2) = res
    assert rel_err(y1, y2.cpu().numpy()) < 1e-5
    assert rel_err(gx1, gx2.cpu().numpy()) < 1e-4
    for a, b in zip(gp1, gp2):
        assert rel_err(a, b.cpu().numpy()) < 1e-3


def test_ss2d_wide_rows_fall_back_to_separate_epilogue_passes():
    """d_inner beyond the fused epilogue's shared-memory tiles (VMamba stage 4: 1536; the kernels take 1664 forward / 832
    backward) must not fail in the middle of backward(): SS2D composes merge + LayerNorm + gate from separate passes.
    Checked against the CPU oracle (oracle/ss2d_ref.py) on a 4 x 5 map."""
    import ceigm_unet_b200 as P
    from oracle import ss2d_ref
    assert P.ops.out_gate_max_D(True) < 1536 <= P.ops.out_gate_max_D(False)
    torch.manual_seed(3)
    m = P.SS2D(d_model=768, d_state=4, ssm_ratio=2.0, k_group=4)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    x = torch.randn(1, 4, 5, 768)
    gy = torch.randn(1, 4, 5, 768)
    xg = x.cuda().requires_grad_(True)
    y = m(xg)
    y.backward(gy.cuda())
    ss2d_ref.use_fast_scan(True)
    try:
        p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xr = x.clone().requires_grad_(True)
        yr = ss2d_ref.ss2d_forward(xr, p, (1, 2, 3, 4))
        yr.backward(gy)
    finally:
        ss2d_ref.use_fast_scan(False)
    assert rel_err(y, yr.detach().numpy()) < 1e-3
    assert rel_err(xg.grad, xr.grad.numpy()) < 1e-3
    for n, q in m.named_parameters():
        assert rel_err(q.grad, p[n].grad.numpy()) < 2e-3, n
    with torch.no_grad():                                   # no_grad: the forward kernel alone covers 1536
        assert rel_err(m(xg), yr.detach().numpy()) < 1e-3


def test_second_device_launches_with_opted_in_shared_memory():
    """cudaFuncSetAttribute is per device: the first launch on cuda:1 in a process that already used cuda:0 must succeed."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ceigm_unet_b200 as P
    from oracle.selective_scan_ref import make_inputs
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        a = make_inputs(2, 64, 256, 16, groups=4, seed=1, device=dev, requires_grad=True)
        o = P.SelectiveScanCore.apply(a["u"], a["delta"], a["A"], a["B"], a["C"], a["D"], a["delta_bias"], True)
        o.backward(a["dout"])
        torch.cuda.synchronize(dev)
        outs.append((o.detach().cpu(), a["u"].grad.cpu()))
    assert torch.allclose(outs[0][0], outs[1][0], atol=1e-5) and torch.allclose(outs[0][1], outs[1][1], atol=1e-5)


@pytest.mark.parametrize("C,H", [(64, 56), (348, 14), (448, 7), (32, 6)])
def test_grouped_layer_equals_the_four_separate_ss2ds(C, H):
    """GroupMambaLayer's grouped pipeline (block-structured projections, one conv, one G = 4 scan launch, one grouped
    epilogue launch) against the same module running its four SS2Ds one after the other (grouped = False)."""
    import copy

    import ceigm_unet_b200 as P
    torch.manual_seed(C + H)
    a = P.GroupMambaLayer(C, C).cuda()
    with torch.no_grad():
        for n, p_ in a.named_parameters():
            if "norm" in n or n == "skip_scale":
                p_.add_(0.1 * torch.randn_like(p_))
    b = copy.deepcopy(a)
    b.grouped = False
    x = torch.randn(2, H * H, C, device="cuda")
    gy = torch.randn(2, H * H, C, device="cuda")
    res = []
    for m in (a, b):
        xi = x.clone().requires_grad_(True)
        y = m(xi, H, H)
        y.backward(gy)
        res.append((y.detach(), xi.grad, {n: p_.grad for n, p_ in m.named_parameters()}))
    (y1, gx1, gp1), (y2, gx2, gp2) = res
    assert rel_err(y1, y2.cpu().numpy()) < 1e-4
    assert rel_err(gx1, gx2.cpu().numpy()) < 1e-4
    for n in gp1:
        assert rel_err(gp1[n], gp2[n].cpu().numpy()) < 5e-4, n


def _tc_calls(monkeypatch):
    """Counts the launches of the two tcgen05 kernels: in_proj (ops.linear_tc) and the fused epilogue + out_proj
    (ops.gate_proj_fwd)."""
    from ceigm_unet_b200 import ops
    calls = []
    real_l, real_g = ops.linear_tc, ops.gate_proj_fwd
    monkeypatch.setattr(ops, "linear_tc", lambda *a, **k: (calls.append("linear_tc"), real_l(*a, **k))[1])
    monkeypatch.setattr(ops, "gate_proj_fwd", lambda *a, **k: (calls.append("gate_proj"), real_g(*a, **k))[1])
    return calls


def test_ss2d_tensor_core_projections_tf32_vs_fp32(monkeypatch):
    """in_proj + chunk + NHWC->NCHW (ss2d.py:504-510) and out_proj (:518) on the tcgen05 kernel (TF32 math, taken when the
    user allows TF32 matmuls as train_synapse.py:21 does) against the same module with full-fp32 library GEMMs: outputs and
    input gradient rel <= 1e-3, parameter gradients <= 2e-3, at the north-star shape (K = 4, D = 192, 56 x 56)."""
    import ceigm_unet_b200 as P
    torch.manual_seed(3)
    m = P.SS2D(d_model=96, d_state=16, ssm_ratio=2.0, k_group=4).cuda()
    x = torch.randn(2, 56, 56, 96, device="cuda")
    dy = torch.randn(2, 56, 56, 96, device="cuda")
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    res = {}
    try:
        torch.backends.cudnn.allow_tf32 = False
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            calls = _tc_calls(monkeypatch) if tf32 else []
            m.zero_grad(set_to_none=True)
            xg = x.clone().requires_grad_(True)
            y = m(xg)
            y.backward(dy)
            res[tf32] = (y.detach().clone(), xg.grad.clone(), {n: p.grad.clone() for n, p in m.named_parameters()})
            if tf32:
                assert calls == ["linear_tc", "linear_tc"], "in_proj and out_proj must both run on the tensor-core kernel"
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    (y0, dx0, g0), (y1, dx1, g1) = res[False], res[True]
    assert rel_err(y1, y0.cpu().numpy()) < 1e-3
    assert rel_err(dx1, dx0.cpu().numpy()) < 1e-3
    for n in g0:
        assert rel_err(g1[n], g0[n].cpu().numpy()) < 2e-3, n


def test_ss2d_tensor_core_projections_bf16_autocast(monkeypatch):
    """Under bf16 autocast (the training configuration) both projections run on the tensor-core kernel with bf16 operands;
    result within 2e-2 of the fp32 module, like the library path."""
    import ceigm_unet_b200 as P
    torch.manual_seed(4)
    m = P.SS2D(d_model=96, d_state=16, ssm_ratio=2.0, k_group=4).cuda()
    x = torch.randn(2, 28, 28, 96, device="cuda", requires_grad=True)
    dy = torch.randn(2, 28, 28, 96, device="cuda")
    y32 = m(x)
    y32.backward(dy)
    ref = (y32.detach().clone(), x.grad.clone(), {n: p.grad.clone() for n, p in m.named_parameters()})
    m.zero_grad(set_to_none=True)
    x.grad = None
    calls = _tc_calls(monkeypatch)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y16 = m(x)
    y16.backward(dy.to(y16.dtype))
    assert calls == ["linear_tc", "linear_tc"] and y16.dtype == torch.bfloat16
    assert rel_err(y16.float(), ref[0].cpu().numpy()) < 2e-2
    assert rel_err(x.grad, ref[1].cpu().numpy()) < 2e-2
    for n, p in m.named_parameters():
        assert rel_err(p.grad.float(), ref[2][n].cpu().numpy()) < 4e-2, n


def test_ss2d_tensor_core_path_cuda_graphs_match_eager():
    """The tcgen05 projections (tensor maps encoded per call, passed by value) and the two-plane convolution kernel under
    graphed(): replays with new input values reproduce the eager module (TF32 matmuls allowed on both sides)."""
    import copy

    import ceigm_unet_b200 as P
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        torch.manual_seed(2)
        eager = P.SS2D(d_model=32, d_state=16, ssm_ratio=2.0, k_group=4).cuda()
        captured = copy.deepcopy(eager)
        x = torch.randn(3, 12, 16, 32, device="cuda", requires_grad=True)
        gy = torch.randn(3, 12, 16, 32, device="cuda")
        g = P.graphed(captured, (x.detach().clone().requires_grad_(True),))
        res = []
        for fn, mod in ((g, captured), (eager, eager)):
            for rep in range(2):
                xin = (x.detach() * (1.0 + 0.5 * rep)).requires_grad_(True)
                for p_ in mod.parameters():
                    p_.grad = None
                y = fn(xin)
                y.backward(gy)
            res.append((y.detach().clone(), xin.grad.clone(), [p_.grad.clone() for p_ in mod.parameters()]))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    (y1, gx1, gp1), (y2, gx2, gp2) = res
    assert rel_err(y1, y2.cpu().numpy()) < 1e-5
    assert rel_err(gx1, gx2.cpu().numpy()) < 1e-4
    for a, b in zip(gp1, gp2):
        assert rel_err(a, b.cpu().numpy()) < 1e-3


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_ss2d_fused_epilogue_and_out_proj_kernel(monkeypatch, mode):
    """SS2D.fuse_out_proj = True: merge + out_norm + gate + out_proj (ss2d.py:486-518) in ONE tcgen05 kernel
    (csrc/gate_proj_tc.cu) — output and every gradient against the same module on the two-kernel path."""
    import ceigm_unet_b200 as P
    torch.manual_seed(5)
    m = P.SS2D(d_model=96, d_state=16, ssm_ratio=2.0, k_group=4).cuda()
    x = torch.randn(2, 28, 28, 96, device="cuda")
    dy = torch.randn(2, 28, 28, 96, device="cuda")
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    res = {}
    try:
        for fused in (False, True):
            m.fuse_out_proj = fused
            calls = _tc_calls(monkeypatch) if fused else []
            m.zero_grad(set_to_none=True)
            xg = x.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
                y = m(xg)
            y.backward(dy.to(y.dtype))
            res[fused] = (y.detach().float().clone(), xg.grad.clone(), {n: p.grad.clone() for n, p in m.named_parameters()})
            if fused:
                assert calls == ["linear_tc", "gate_proj"]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
        m.fuse_out_proj = False
    tol = 1e-3 if mode == "tf32" else 2e-2
    (y0, dx0, g0), (y1, dx1, g1) = res[False], res[True]
    assert rel_err(y1, y0.cpu().numpy()) < tol
    assert rel_err(dx1, dx0.cpu().numpy()) < tol
    for n in g0:
        assert rel_err(g1[n].float(), g0[n].float().cpu().numpy()) < 2 * tol, n


@pytest.mark.parametrize("C,amp", [(64, False), (448, False), (348, True), (600, False)])
def test_layernorm_rows_module_equals_nn_layernorm(C, amp):
    """ceigm_unet_b200.LayerNormRows (re-classed nn.LayerNorm, csrc/layernorm.cu; wider than 512 channels: nn.LayerNorm's own
    forward) against nn.LayerNorm: output, input and parameter gradients; under bf16 autocast both return fp32."""
    import ceigm_unet_b200 as P
    torch.manual_seed(C)
    ref = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        ref.weight.add_(0.1 * torch.randn_like(ref.weight)); ref.bias.add_(0.1 * torch.randn_like(ref.bias))
    ours = torch.nn.LayerNorm(C).cuda()
    ours.load_state_dict(ref.state_dict())
    ours.__class__ = P.LayerNormRows
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    x = torch.randn(3, 50, C, device="cuda")
    if amp:
        x = x.bfloat16()
    dy = torch.randn(3, 50, C, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        ya, yb = ref(xa), ours(xb)
    assert ya.dtype == yb.dtype
    (ya * dy).sum().backward(); (yb * dy).sum().backward()
    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
    tol = 2e-2 if amp else 1e-4
    assert rel(yb, ya) < 1e-4 and rel(xb.grad, xa.grad) < tol
    assert rel(ours.weight.grad, ref.weight.grad) < 1e-3 and rel(ours.bias.grad, ref.bias.grad) < 1e-3
