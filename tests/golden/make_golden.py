"""Generates tests/golden/*.npz by running the UNMODIFIED reference python from /root/reference
(build container only; see oracle/ref_loader.py for the three shims). Re-run with
    python tests/golden/make_golden.py
The committed fixtures are what tests/test_oracle_golden.py and the `-m gpu` parity tests compare to.
Reference entry points exercised:
  kernels/selective_scan/test_selective_scan.py:168-234   selective_scan_ref  (+ autograd for grads)
  model/gm/csms6s.py:11-206                               CrossScan/CrossMerge[_1.._4]
  model/gm/ss2d.py:521-556                                SS2D (k_group=1, d_state=1)
  model/vmamba/vmamba.py:992                              SS2D (k_group=4, d_state=16, forward_type v2)
  model/gm/groupmamba.py:85-159                           GroupMambaLayer
  model/gm/groupmamba.py:54-83, custom_mlp.py:313-368     PVT2FFN, custom_ffn (+ their initialisation under a fixed seed)
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

from oracle import ref_loader as RL                      # noqa: E402
from oracle.selective_scan_ref import make_inputs        # noqa: E402

# name: (batch, dim, L, N, G, has_D, has_bias, softplus, delta_scale, squeeze_BC)
SCAN_CASES = {
    "scan_n1_g1":        (2, 8, 64, 1, 1, True, True, True, 1.0, False),
    "scan_n4_g2_odd":    (2, 8, 65, 4, 2, True, True, True, 1.0, False),
    "scan_n16_g4":       (2, 12, 37, 16, 4, True, True, True, 1.0, False),
    "scan_n16_long":     (1, 4, 300, 16, 1, True, True, True, 1.0, False),
    "scan_n16_noD":      (2, 4, 33, 16, 2, False, True, True, 1.0, False),
    "scan_n16_nobias":   (2, 4, 33, 16, 2, True, False, True, 1.0, False),
    "scan_n16_nosp":     (2, 4, 33, 16, 2, True, True, False, 1.0, False),
    "scan_n2_thresh20":  (2, 6, 48, 2, 1, True, True, True, 60.0, False),   # delta+bias crosses softplus threshold 20
    "scan_n3_bc3d":      (2, 5, 29, 3, 1, True, True, True, 1.0, True),     # B, C given as (b, n, L)
    "scan_n1_l49":       (2, 16, 49, 1, 1, True, True, True, 1.0, False),   # live-model stage-4 length
}


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def gen_scan(ref_fn):
    for i, (name, (nb, nd, L, N, G, hD, hb, sp, dscale, sq)) in enumerate(SCAN_CASES.items()):
        inp = make_inputs(nb, nd, L, N, groups=G, has_D=hD, has_delta_bias=hb, seed=100 + i)
        inp["delta"] = inp["delta"] * dscale
        if sq:
            inp["B"], inp["C"] = inp["B"][:, 0].contiguous(), inp["C"][:, 0].contiguous()
        leaves = {k: (v.clone().requires_grad_(True) if v is not None else None)
                  for k, v in inp.items() if k != "dout"}
        out, last = ref_fn(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"],
                           None, leaves["delta_bias"], sp, True)
        out.backward(inp["dout"])
        rec = {"in_" + k: _np(v) for k, v in inp.items() if v is not None}
        rec.update(out=_np(out), last_state=_np(last), softplus=np.array(sp))
        for k, v in leaves.items():
            if v is not None:
                rec["grad_" + k] = _np(v.grad)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        print(name, "out", tuple(out.shape))


def gen_cross(cs):
    rec = {}
    for (H, W) in [(6, 6), (4, 6)]:
        x = torch.arange(2 * 3 * H * W, dtype=torch.float32).view(2, 3, H, W)
        tag = f"{H}x{W}"
        rec[f"x_{tag}"] = _np(x)
        rec[f"scanK_{tag}"] = _np(cs.CrossScan.apply(x))
        ys = torch.arange(2 * 4 * 3 * H * W, dtype=torch.float32).view(2, 4, 3, H, W) % 97
        rec[f"ysK_{tag}"] = _np(ys)
        rec[f"mergeK_{tag}"] = _np(cs.CrossMerge.apply(ys))
        # backward of the K=4 pair (the reference's hand-written backward; correct for any H, W)
        xg = x.clone().requires_grad_(True)
        o = cs.CrossScan.apply(xg)
        w = (torch.arange(o.numel(), dtype=torch.float32).view_as(o) % 13) - 6
        (o * w).sum().backward()
        rec[f"scanK_bwd_w_{tag}"], rec[f"scanK_bwd_{tag}"] = _np(w), _np(xg.grad)
        yg = ys.clone().requires_grad_(True)
        o = cs.CrossMerge.apply(yg)
        w = (torch.arange(o.numel(), dtype=torch.float32).view_as(o) % 11) - 5
        (o * w).sum().backward()
        rec[f"mergeK_bwd_w_{tag}"], rec[f"mergeK_bwd_{tag}"] = _np(w), _np(yg.grad)
        for k in (1, 2, 3, 4):
            S, M = getattr(cs, f"CrossScan_{k}"), getattr(cs, f"CrossMerge_{k}")
            rec[f"scan{k}_{tag}"] = _np(S.apply(x))
            rec[f"merge{k}_{tag}"] = _np(M.apply(ys[:, k - 1:k].contiguous()))
            if H == W:   # the reference's _2/_4 backward is only a true adjoint on square maps
                xg = x.clone().requires_grad_(True)
                o = S.apply(xg)
                w = (torch.arange(o.numel(), dtype=torch.float32).view_as(o) % 7) - 3
                (o * w).sum().backward()
                rec[f"scan{k}_bwd_w_{tag}"], rec[f"scan{k}_bwd_{tag}"] = _np(w), _np(xg.grad)
                yg = ys[:, k - 1:k].clone().requires_grad_(True)
                o = M.apply(yg)
                w = (torch.arange(o.numel(), dtype=torch.float32).view_as(o) % 5) - 2
                (o * w).sum().backward()
                rec[f"merge{k}_bwd_w_{tag}"], rec[f"merge{k}_bwd_{tag}"] = _np(w), _np(yg.grad)
    np.savez_compressed(os.path.join(HERE, "cross_scan_merge.npz"), **rec)
    print("cross_scan_merge", len(rec), "arrays")


def _perturb(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in module.named_parameters():
            if n.endswith("A_logs"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            else:
                p.add_(0.05 * torch.randn(p.shape, generator=g))


def _module_record(mod, x, run):
    xg = x.clone().requires_grad_(True)
    y = run(xg)
    w = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
    (y * w).sum().backward()
    rec = dict(x=_np(x), y=_np(y), dy=_np(w), dx=_np(xg.grad))
    for n, p in mod.state_dict().items():
        rec["sd." + n] = _np(p)
    for n, p in mod.named_parameters():
        rec["grad." + n] = _np(p.grad)
    return rec


def gen_ss2d_gm(gm):
    cs = gm["csms6s"]
    for k in (1, 2, 3, 4):
        torch.manual_seed(10 + k)
        m = gm["ss2d"].SS2D(d_model=8, d_state=1, ssm_ratio=1, d_conv=3)
        _perturb(m, 20 + k)
        x = torch.randn(2, 6, 6, 8, generator=torch.Generator().manual_seed(30 + k))
        rec = _module_record(m, x, lambda t: m(t, CrossScan=getattr(cs, f"CrossScan_{k}"),
                                               CrossMerge=getattr(cs, f"CrossMerge_{k}")))
        np.savez_compressed(os.path.join(HERE, f"ss2d_gm_dir{k}.npz"), **rec)
        print(f"ss2d_gm_dir{k}", rec["y"].shape)


def gen_ss2d_vm(vm):
    torch.manual_seed(5)
    m = vm["vmamba"].SS2D(d_model=8, d_state=16, ssm_ratio=2.0, forward_type="v2")
    _perturb(m, 6)
    x = torch.randn(2, 5, 6, 8, generator=torch.Generator().manual_seed(8))     # non-square on purpose
    rec = _module_record(m, x, lambda t: m(t))
    np.savez_compressed(os.path.join(HERE, "ss2d_vm_k4_n16.npz"), **rec)
    print("ss2d_vm_k4_n16", rec["y"].shape)


def gen_group_layer(gm):
    torch.manual_seed(3)
    m = gm["groupmamba"].GroupMambaLayer(32, 32)
    _perturb(m, 4)
    x = torch.randn(2, 36, 32, generator=torch.Generator().manual_seed(9))
    rec = _module_record(m, x, lambda t: m(t, 6, 6))
    np.savez_compressed(os.path.join(HERE, "group_mamba_layer.npz"), **rec)
    print("group_mamba_layer", rec["y"].shape)


def gen_init(gm, vm):
    """Parameters right after construction under a fixed seed: pins the RNG consumption order of the constructors
    (mamba_init, ss2d.py:154-209; __initv2__, ss2d.py:294-335)."""
    rec = {}
    torch.manual_seed(1234)
    m = gm["ss2d"].SS2D(d_model=32, d_state=1, ssm_ratio=1, d_conv=3)
    rec.update({"gm." + k: _np(v) for k, v in m.state_dict().items()})
    torch.manual_seed(1234)
    m = vm["vmamba"].SS2D(d_model=16, d_state=16, ssm_ratio=2.0, forward_type="v2")
    rec.update({"vm." + k: _np(v) for k, v in m.state_dict().items()})
    torch.manual_seed(1234)
    m = gm["groupmamba"].GroupMambaLayer(64, 64)
    rec.update({"layer." + k: _np(v) for k, v in m.state_dict().items()})
    np.savez_compressed(os.path.join(HERE, "init_seed1234.npz"), **rec)
    print("init_seed1234", len(rec), "arrays")


def gen_ffn(gm):
    """The feed-forward blocks next to the SS2D branch, non-square maps on purpose."""
    torch.manual_seed(11)
    m = gm["groupmamba"].PVT2FFN(16, 64)
    init = {"init." + k: _np(v).copy() for k, v in m.state_dict().items()}
    _perturb(m, 12)
    x = torch.randn(2, 30, 16, generator=torch.Generator().manual_seed(13))
    rec = _module_record(m, x, lambda t: m(t, 6, 5))
    rec.update(init)
    np.savez_compressed(os.path.join(HERE, "ffn_pvt2.npz"), **rec)
    print("ffn_pvt2", rec["y"].shape)
    torch.manual_seed(14)
    m = gm["custom_mlp"].custom_ffn(12, 48)              # gc = 6: segments 30 | 6 | 6 | 6
    init = {"init." + k: _np(v).copy() for k, v in m.state_dict().items()}
    _perturb(m, 15)
    x = torch.randn(2, 56, 12, generator=torch.Generator().manual_seed(16))
    rec = _module_record(m, x, lambda t: m(t, 8, 7))
    rec.update(init)
    np.savez_compressed(os.path.join(HERE, "ffn_custom.npz"), **rec)
    print("ffn_custom", rec["y"].shape)


if __name__ == "__main__":
    if not RL.available():
        raise SystemExit("reference tree not present: golden vectors can only be regenerated in the build container")
    gen_scan(RL.load_selective_scan_ref())
    gm = RL.load_gm()
    gen_cross(gm["csms6s"])
    gen_ss2d_gm(gm)
    vm = RL.load_vmamba()
    gen_ss2d_vm(vm)
    gen_group_layer(gm)
    gen_init(gm, vm)
    gen_ffn(gm)
