"""Loads the UNMODIFIED reference GM-UNet (`baseline/_ref/gm-unet`, installed by harness/install_ref.py; in the build
container `/root/reference/gm-unet` is the fallback) and runs it on top of this repo's drop-in (boundary b4).

No reference file is edited. What the harness does instead:
  * `harness/shims` goes on sys.path: stand-ins for the absent timm / monai / calflops imports;
  * `ceigm_unet_b200.install_dropin()` registers `selective_scan_cuda_core` / `selective_scan_cuda_oflex`, which the
    reference imports in try/except at model/gm/csms6s.py:209-220 — level-1 drop-in: every scan goes through libss2d_b200.so;
  * `model.EMCAD22nn = model.EMCAD22n` (model/__init__.py:9 binds EMCAD22n, :29 reads EMCAD22nn -> NameError);
  * level 2 (`fused=True`): `model.gm.groupmamba.GroupMambaLayer` is rebound to `ceigm_unet_b200.GroupMambaLayer` before the
    model is built (Block_mamba looks the name up at construction, groupmamba.py:203) — same state_dict, fused kernels;
    likewise `PVT2FFN` (groupmamba.py:204) and the decoder's `custom_ffn` (custom_module.py:51) -> the channels-last
    depthwise-stack kernels (csrc/ffn_dw.cu).

`scan="cpu_ref"` is for the CHECKER side of tests and the CPU baseline of bench.py only: SelectiveScanCore is rebound to the
oracle's restatement of the reference's PyTorch scan (the reference has no CPU scan, csms6s.py:352).
"""
from __future__ import annotations

import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "harness", "shims")
_CANDIDATES = (os.path.join(ROOT, "baseline", "_ref", "gm-unet"), "/root/reference/gm-unet")
REF_EXT = os.path.join(ROOT, "baseline", "_ref", "ext", "ref_selective_scan_cuda_core.so")


def ref_root():
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "model", "__init__.py")):
            return c
    return None


def available() -> bool:
    return ref_root() is not None


def _paths():
    root = ref_root()
    if root is None:
        raise RuntimeError("reference tree not found: run `python harness/install_ref.py` in the build container "
                           "(it populates the git-ignored baseline/_ref/, which travels to the GPU box)")
    for p in (SHIMS, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    return root


def _purge():
    """Forget previously imported reference modules so that a different binding (drop-in / cpu_ref / fused) takes effect."""
    for name in list(sys.modules):
        if name == "model" or name.startswith("model.") or name in ("loss", "utils", "eval"):
            del sys.modules[name]


def load_reference(scan: str = "dropin", fused: bool = False):
    """-> the reference's top-level `model` package, ready for `model.build_model(in_channels=3, num_classes=9)`.

    scan = "dropin":  the reference's SelectiveScanCore calls this repo's extension modules (CUDA only, no fallback).
    scan = "cpu_ref": SelectiveScanCore := oracle restatement of selective_scan_ref (checker / CPU baseline only).
    scan = "cpu_fast": the same through the C oracle (fast analytic backward; checker only).
    fused = True:     GroupMambaLayer := ceigm_unet_b200.GroupMambaLayer (level-2 drop-in; implies scan="dropin")."""
    _paths()
    _purge()
    if scan == "dropin":
        import ceigm_unet_b200 as pkg
        pkg.install_dropin()
    model = importlib.import_module("model")
    model.EMCAD22nn = model.EMCAD22n
    ss2d = importlib.import_module("model.gm.ss2d")
    gmb = importlib.import_module("model.gm.groupmamba")
    if scan in ("cpu_ref", "cpu_fast"):
        if scan == "cpu_ref":
            from oracle.selective_scan_ref import selective_scan_ref
        else:      # same function, C oracle behind an autograd.Function (oracle/fast_scan.py)
            from oracle.fast_scan import selective_scan_fast as selective_scan_ref

        class SelectiveScanRefCPU:
            @staticmethod
            def apply(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
                return selective_scan_ref(u, delta, A, B, C, D, None, delta_bias, delta_softplus)
        ss2d.SelectiveScanCore = SelectiveScanRefCPU         # looked up in SS2D.__initv2__ (ss2d.py:287)
    elif scan != "dropin":
        raise ValueError(scan)
    if fused:
        import ceigm_unet_b200 as pkg
        gmb.GroupMambaLayer = pkg.GroupMambaLayer
        # the FFNs next to the SS2D branch: Block_mamba reads PVT2FFN from groupmamba's globals
        # (groupmamba.py:204), the decoder's `cm` passes custom_module's `custom_ffn` (custom_module.py:51)
        gmb.PVT2FFN = pkg.PVT2FFN
        importlib.import_module("model.gm.custom_module").custom_ffn = pkg.custom_ffn
    return model


def reference_modules():
    """The reference's gm sub-modules after `load_reference` -> (csms6s, ss2d, groupmamba)."""
    return tuple(importlib.import_module("model.gm." + n) for n in ("csms6s", "ss2d", "groupmamba"))


def load_losses():
    _paths()
    return importlib.import_module("loss")


def load_ref_cuda_ext():
    """The reference's own CUDA extension recompiled for sm_100 (baseline column only) or None if it was not built."""
    if not os.path.exists(REF_EXT):
        return None
    import importlib.util

    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location("ref_selective_scan_cuda_core", REF_EXT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reclass_layernorms(net):
    """Level 2 also covers the LayerNorms of the GroupMamba encoder (Block_mamba.norm2 in front of the FFN, the patch-embedding
    and stage norms: groupmamba.py:200, 225, 231): every plain nn.LayerNorm over <= 512 channels is RE-CLASSED in place to
    ceigm_unet_b200.LayerNormRows (this repo's row kernel, csrc/layernorm.cu) — same parameters, same state_dict. Returns the
    number of modules switched."""
    import torch
    import ceigm_unet_b200 as pkg
    n = 0
    for mod in net.modules():
        if type(mod) is torch.nn.LayerNorm and len(mod.normalized_shape) == 1 and mod.elementwise_affine \
                and mod.normalized_shape[0] <= pkg.ops.LN_MAX_C:
            mod.__class__ = pkg.LayerNormRows
            n += 1
    return n
