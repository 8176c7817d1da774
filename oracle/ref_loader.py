"""ORACLE tooling (test infrastructure only). Loads the UNMODIFIED reference python from
/root/reference (exists in the build container only — never on the GPU box) so that
tests/golden/make_golden.py can record what the reference itself computes.

No reference file is edited or copied. Shims:
  * `torch.Tensor.embed_dim = torch.Tensor.dim` while `selective_scan_ref` runs (typo at
    kernels/selective_scan/test_selective_scan.py:191-192);
  * a stub `timm` (DropPath, trunc_normal_, register_model, _cfg) because timm is not installed;
  * `SelectiveScanCore` rebound to a class whose `.apply` calls the reference's own
    `selective_scan_ref` (the reference has no CPU scan: model/gm/csms6s.py:352 calls the CUDA ext).
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = "/root/reference/gm-unet"


def available() -> bool:
    return os.path.isdir(REF_ROOT)


def load_selective_scan_ref():
    """Returns the reference's `selective_scan_ref` function object, compiled from its own source."""
    from einops import rearrange, repeat
    path = os.path.join(REF_ROOT, "kernels/selective_scan/test_selective_scan.py")
    tree = ast.parse(open(path).read(), filename=path)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "selective_scan_ref"][0]
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = dict(torch=torch, F=F, rearrange=rearrange, repeat=repeat)
    exec(compile(mod, path, "exec"), ns)
    raw = ns["selective_scan_ref"]

    def selective_scan_ref(*a, **k):
        had = hasattr(torch.Tensor, "embed_dim")
        torch.Tensor.embed_dim = torch.Tensor.dim      # the typo shim, active only during the call
        try:
            return raw(*a, **k)
        finally:
            if not had:
                del torch.Tensor.embed_dim
    return selective_scan_ref


def _install_timm_stub():
    if "timm" in sys.modules:
        return

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            return x * mask / keep

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    registry = types.ModuleType("timm.models.registry")
    vit = types.ModuleType("timm.models.vision_transformer")
    helpers = types.ModuleType("timm.models.helpers")
    layers.DropPath, layers.trunc_normal_, layers.trunc_normal_tf_ = DropPath, trunc_normal_, trunc_normal_
    registry.register_model = lambda fn: fn
    vit._cfg = lambda **kw: dict(kw)

    def named_apply(fn, module, name="", depth_first=True, include_root=False):
        for cn, cm in module.named_children():
            named_apply(fn, cm, ".".join((name, cn)) if name else cn, depth_first, True)
        if include_root:
            fn(module=module, name=name)
        return module
    helpers.named_apply = named_apply
    timm.models, models.layers, models.registry = models, layers, registry
    models.vision_transformer, models.helpers = vit, helpers
    for m in (timm, models, layers, registry, vit, helpers):
        sys.modules[m.__name__] = m


def _load_pkg(alias: str, rel_dir: str, names):
    """Import files of a reference sub-package under a private alias (skips the heavy model/__init__.py)."""
    pkg = types.ModuleType(alias)
    pkg.__path__ = [os.path.join(REF_ROOT, rel_dir)]
    sys.modules[alias] = pkg
    out = {}
    for n in names:
        spec = importlib.util.spec_from_file_location(f"{alias}.{n}", os.path.join(REF_ROOT, rel_dir, n + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        out[n] = mod
    return out


def make_cpu_scan_function(selective_scan_ref):
    """Class with the `SelectiveScanCore.apply` signature (model/gm/csms6s.py:350) backed by the
    reference's PyTorch scan; differentiable by autograd."""
    class SelectiveScanRefCPU:
        @staticmethod
        def apply(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
            return selective_scan_ref(u, delta, A, B, C, D, None, delta_bias, delta_softplus)
    return SelectiveScanRefCPU


_CACHE = {}


def load_gm():
    """-> dict(csms6s, ss2d, groupmamba) reference modules (model/gm/*), CPU scan rebound."""
    if "gm" not in _CACHE:
        _install_timm_stub()
        mods = _load_pkg("_ref_gm", "model/gm", ["csms6s", "ss2d", "custom_mlp", "groupmamba"])
        cpu = make_cpu_scan_function(load_selective_scan_ref())
        mods["ss2d"].SelectiveScanCore = cpu          # looked up at SS2D.__init__ time (ss2d.py:287)
        _CACHE["gm"] = mods
    return _CACHE["gm"]


def load_vmamba():
    """-> dict(csms6s, vmamba) reference modules (model/vmamba/*), CPU scan rebound."""
    if "vm" not in _CACHE:
        _install_timm_stub()
        mods = _load_pkg("_ref_vm", "model/vmamba", ["csms6s", "csm_triton", "vmamba"])
        cpu = make_cpu_scan_function(load_selective_scan_ref())
        mods["vmamba"].SelectiveScanCore = cpu
        _CACHE["vm"] = mods
    return _CACHE["vm"]
