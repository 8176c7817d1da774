"""CPU tests: every oracle restatement against the golden vectors recorded from the reference itself
(tests/golden/make_golden.py). These pin the oracle; the GPU parity tests then compare CUDA vs oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, cross_scan, ffn_ref, scan_analytic, ss2d_ref
from oracle.selective_scan_ref import selective_scan_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCAN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "scan_*.npz")))
GRADS = [("grad_u", "du"), ("grad_delta", "ddelta"), ("grad_A", "dA"), ("grad_B", "dB"), ("grad_C", "dC"),
         ("grad_D", "dD"), ("grad_delta_bias", "ddelta_bias")]


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def test_golden_present():
    assert len(SCAN_FILES) >= 10


@pytest.mark.parametrize("path", SCAN_FILES, ids=lambda p: os.path.basename(p)[:-4])
def test_torch_restatement_matches_reference(path):
    g = load(path)
    t = {k[3:]: torch.from_numpy(v).requires_grad_(k != "in_dout") for k, v in g.items() if k.startswith("in_")}
    out, last = selective_scan_ref(t["u"], t["delta"], t["A"], t["B"], t["C"], t.get("D"), None,
                                   t.get("delta_bias"), bool(g["softplus"]), True)
    out.backward(t["dout"])
    assert rel_err(out.detach().numpy(), g["out"]) < 1e-6
    assert rel_err(last.detach().numpy(), g["last_state"]) < 1e-6
    for gk, _ in GRADS:
        if gk in g:
            assert rel_err(t[gk[5:]].grad.numpy(), g[gk]) < 1e-5, gk


@pytest.mark.parametrize("path", SCAN_FILES, ids=lambda p: os.path.basename(p)[:-4])
def test_analytic_backward_matches_reference(path):
    g = load(path)
    t = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("in_")}
    r = scan_analytic.scan_fwd_bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t.get("D"), t.get("delta_bias"),
                                   bool(g["softplus"]), t["dout"])
    assert rel_err(r["out"].numpy(), g["out"]) < 2e-6
    assert rel_err(r["last_state"].numpy(), g["last_state"]) < 2e-6
    for gk, ak in GRADS:
        if gk in g:
            assert rel_err(r[ak].numpy(), g[gk]) < 2e-5, gk      # reference grads are fp32 autograd


@pytest.mark.parametrize("acc,tol", [("f64", 2e-5), ("f32", 2e-4)])
@pytest.mark.parametrize("path", SCAN_FILES, ids=lambda p: os.path.basename(p)[:-4])
def test_c_oracle_matches_reference(path, acc, tol):
    g = load(path)
    i = {k[3:]: v for k, v in g.items() if k.startswith("in_")}
    Bm, Cm = i["B"], i["C"]
    if Bm.ndim == 3:
        Bm, Cm = Bm[:, None], Cm[:, None]
    out, last = c_oracle.scan_fwd(i["u"], i["delta"], i["A"], Bm, Cm, i.get("D"), i.get("delta_bias"),
                                  bool(g["softplus"]), acc=acc)
    assert rel_err(out, g["out"]) < tol
    assert rel_err(last, g["last_state"]) < tol
    r = c_oracle.scan_bwd(i["u"], i["delta"], i["A"], Bm, Cm, i.get("D"), i.get("delta_bias"), i["dout"],
                          bool(g["softplus"]), acc=acc)
    for gk, ak in GRADS:
        if gk in g:
            got = r[ak]
            if ak in ("dB", "dC") and g[gk].ndim == 3:
                got = got[:, 0]
            assert rel_err(got, g[gk]) < tol, gk


def test_cross_scan_merge_bit_exact():
    g = load(os.path.join(GOLDEN, "cross_scan_merge.npz"))
    for tag in ("6x6", "4x6"):
        x, ys = torch.from_numpy(g[f"x_{tag}"]), torch.from_numpy(g[f"ysK_{tag}"])
        assert np.array_equal(cross_scan.cross_scan4(x).numpy(), g[f"scanK_{tag}"])
        assert np.array_equal(cross_scan.cross_merge4(ys).numpy(), g[f"mergeK_{tag}"])
        for k in (1, 2, 3, 4):
            assert np.array_equal(cross_scan.cross_scan_k(x, k).numpy(), g[f"scan{k}_{tag}"])
            assert np.array_equal(cross_scan.cross_merge_k(ys[:, k - 1:k], k).numpy(), g[f"merge{k}_{tag}"])
        # adjoints (autograd of the restated forward) vs the reference's hand-written backward
        xg = x.clone().requires_grad_(True)
        (cross_scan.cross_scan4(xg) * torch.from_numpy(g[f"scanK_bwd_w_{tag}"])).sum().backward()
        assert np.array_equal(xg.grad.numpy(), g[f"scanK_bwd_{tag}"])
        yg = ys.clone().requires_grad_(True)
        (cross_scan.cross_merge4(yg) * torch.from_numpy(g[f"mergeK_bwd_w_{tag}"])).sum().backward()
        assert np.array_equal(yg.grad.numpy(), g[f"mergeK_bwd_{tag}"])
    for k in (1, 2, 3, 4):      # single-direction adjoints: square map only (reference _2/_4 backward bug on H != W)
        x = torch.from_numpy(g["x_6x6"]).requires_grad_(True)
        (cross_scan.cross_scan_k(x, k) * torch.from_numpy(g[f"scan{k}_bwd_w_6x6"])).sum().backward()
        assert np.array_equal(x.grad.numpy(), g[f"scan{k}_bwd_6x6"])
        ys = torch.from_numpy(g["ysK_6x6"][:, k - 1:k]).requires_grad_(True)
        (cross_scan.cross_merge_k(ys, k) * torch.from_numpy(g[f"merge{k}_bwd_w_6x6"])).sum().backward()
        assert np.array_equal(ys.grad.numpy(), g[f"merge{k}_bwd_6x6"])


def _module_case(path, fn):
    g = load(path)
    p = {k[3:]: torch.from_numpy(v).requires_grad_(True) for k, v in g.items() if k.startswith("sd.")}
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = fn(x, p)
    (y * torch.from_numpy(g["dy"])).sum().backward()
    assert rel_err(y.detach().numpy(), g["y"]) < 1e-5
    assert rel_err(x.grad.numpy(), g["dx"]) < 1e-4
    for k, v in g.items():
        if k.startswith("grad."):
            assert rel_err(p[k[5:]].grad.numpy(), v) < 2e-4, k


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_ss2d_gm_restatement(k):
    _module_case(os.path.join(GOLDEN, f"ss2d_gm_dir{k}.npz"), lambda x, p: ss2d_ref.ss2d_forward(x, p, (k,)))


def test_ss2d_vm_restatement():
    _module_case(os.path.join(GOLDEN, "ss2d_vm_k4_n16.npz"), lambda x, p: ss2d_ref.ss2d_forward(x, p, (1, 2, 3, 4)))


def test_group_layer_restatement():
    _module_case(os.path.join(GOLDEN, "group_mamba_layer.npz"), lambda x, p: ss2d_ref.group_layer(x, p, 6, 6))


def test_ffn_restatements():
    """oracle/ffn_ref.py vs the unmodified PVT2FFN / custom_ffn (outputs, input and parameter gradients)."""
    _module_case(os.path.join(GOLDEN, "ffn_pvt2.npz"), lambda x, p: ffn_ref.pvt2_ffn(x, p, 6, 5))
    _module_case(os.path.join(GOLDEN, "ffn_custom.npz"), lambda x, p: ffn_ref.custom_ffn(x, p, 8, 7))


def test_ffn_modules_same_keys_and_init():
    """ceigm_unet_b200.PVT2FFN / custom_ffn: the reference's state_dict keys and, under the same seed, bit-identical
    parameters (constructor RNG order + _init_weights: groupmamba.py:55-76, custom_mlp.py:339-361)."""
    import ceigm_unet_b200 as pkg
    for name, ctor, seed in (("ffn_pvt2.npz", lambda: pkg.PVT2FFN(16, 64), 11), ("ffn_custom.npz", lambda: pkg.custom_ffn(12, 48), 14)):
        g = load(os.path.join(GOLDEN, name))
        torch.manual_seed(seed)
        m = ctor()
        sd = m.state_dict()
        assert list(sd.keys()) == [k[5:] for k in g if k.startswith("init.")]
        for k, v in sd.items():
            assert np.array_equal(v.numpy(), g["init." + k]), k


def test_fast_scan_matches_ref():
    """oracle/fast_scan.py (C oracle behind an autograd.Function) == autograd through the restated reference loop."""
    from oracle.fast_scan import selective_scan_fast
    from oracle.selective_scan_ref import make_inputs
    for (b, d, L, n, g) in ((2, 8, 37, 1, 1), (2, 12, 64, 16, 4), (1, 6, 50, 4, 2)):
        a = make_inputs(b, d, L, n, groups=g, seed=3, requires_grad=True)
        c = {k: (v.detach().clone().requires_grad_(v.requires_grad) if v is not None else None) for k, v in a.items()}
        o1 = selective_scan_ref(a["u"], a["delta"], a["A"], a["B"], a["C"], a["D"], None, a["delta_bias"], True)
        o2 = selective_scan_fast(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], None, c["delta_bias"], True)
        o1.backward(a["dout"]); o2.backward(c["dout"])
        assert rel_err(o2.detach().numpy(), o1.detach().numpy()) < 1e-5
        for k in ("u", "delta", "A", "B", "C", "D", "delta_bias"):
            assert rel_err(c[k].grad.numpy(), a[k].grad.numpy()) < 2e-5, k


def test_group_layer_oracle_matches_reference_module_live_shape():
    """The restated GroupMambaLayer with the fast scan against the UNMODIFIED reference module (python scan) at a live-sized
    channel count (C = 64 as in stage 1, a 12 x 12 map to keep the python loop short). Needs the reference tree."""
    from harness import refmodel
    if not refmodel.available():
        pytest.skip("reference tree not installed (baseline/_ref)")
    refmodel.load_reference(scan="cpu_ref")
    _, _, gmb = refmodel.reference_modules()
    torch.manual_seed(5)
    layer = gmb.GroupMambaLayer(64, 64)
    x = torch.randn(2, 144, 64, requires_grad=True)
    y = layer(x, 12, 12)
    dy = torch.randn_like(y)
    y.backward(dy)
    ss2d_ref.use_fast_scan(True)
    try:
        p = {k: v.detach().clone().requires_grad_(True) for k, v in layer.state_dict().items()}
        x2 = x.detach().clone().requires_grad_(True)
        y2 = ss2d_ref.group_layer(x2, p, 12, 12)
        y2.backward(dy)
    finally:
        ss2d_ref.use_fast_scan(False)
    assert rel_err(y2.detach().numpy(), y.detach().numpy()) < 1e-5
    assert rel_err(x2.grad.numpy(), x.grad.numpy()) < 1e-4
    for n, prm in layer.named_parameters():
        assert rel_err(p[n].grad.numpy(), prm.grad.numpy()) < 1e-4, n
