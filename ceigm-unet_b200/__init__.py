"""ceigm-unet_b200 — B200-native (sm_100a) implementation of GM-UNet's SS2D selective-scan hot path.

Layers:
  b1  dropin/selective_scan_cuda_core.py, dropin/selective_scan_cuda_oflex.py — the reference extension modules'
      `fwd` / `bwd` functions; `install_dropin()` registers them in sys.modules so that the reference's own
      `model/gm/csms6s.py` picks them up unchanged.
  b2  functional.py — SelectiveScanCore/Oflex, CrossScan[_k], CrossMerge[_k] autograd Functions.
  b3  modules.py — SS2D, GroupMambaLayer with the reference's state_dict.
Underneath: ops.py -> _lib.py (ctypes) -> libss2d_b200.so (csrc/*.cu, C ABI in include/ss2d_b200.h).
There is no CPU or PyTorch fallback: without the built library every operator raises RuntimeError.
"""
from . import _lib, dist, functional, modules, ops                 # noqa: F401
from ._lib import build, launch_count, version                    # noqa: F401
from .functional import (CrossMerge, CrossMerge_1, CrossMerge_2, CrossMerge_3, CrossMerge_4, CrossScan,   # noqa: F401
                         CrossScan_1, CrossScan_2, CrossScan_3, CrossScan_4, SelectiveScanCore, SelectiveScanOflex)
from .modules import SS2D, GroupMambaLayer, mamba_init             # noqa: F401


def install_dropin() -> None:
    """Make `import selective_scan_cuda_core` / `selective_scan_cuda_oflex` resolve to this package's modules
    (the reference imports them at module import time inside try/except: model/gm/csms6s.py:209-220)."""
    import sys

    from .dropin import selective_scan_cuda_core, selective_scan_cuda_oflex
    sys.modules["selective_scan_cuda_core"] = selective_scan_cuda_core
    sys.modules["selective_scan_cuda_oflex"] = selective_scan_cuda_oflex
