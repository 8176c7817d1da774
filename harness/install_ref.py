#!/usr/bin/env python
"""Installs the UNMODIFIED reference under the git-ignored `baseline/_ref/` so that it travels to the GPU box:

  baseline/_ref/gm-unet/            a verbatim copy of /root/reference/gm-unet (python model, scripts, kernel sources)
  baseline/_ref/ext/ref_selective_scan_cuda_core.so
                                    the reference's own `cus/` CUDA extension (kernels/selective_scan/csrc/selective_scan/cus,
                                    the sources and nvcc flags of kernels/selective_scan/setup.py:75-131) compiled for sm_100 —
                                    setup.py:62-65 only emits sm_70/80/90, which cannot run on a B200. Built under another
                                    module name so that it can be loaded next to this repo's drop-in of the same name. It is
                                    the "same-box GPU baseline" column of bench.py and never part of the product path.

Run in the build container (where /root/reference exists):  python harness/install_ref.py [--no-ext]
Nothing here is imported by ceigm_unet_b200.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/gm-unet"
DST = os.path.join(ROOT, "baseline", "_ref")
EXT_NAME = "ref_selective_scan_cuda_core"


def copy_tree() -> str:
    dst = os.path.join(DST, "gm-unet")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(SRC, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".ipynb_checkpoints"))
    return dst


def build_ext(verbose: bool = False) -> str:
    """The reference extension compiled from the sources where they lie, for sm_100 (PTX + SASS)."""
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension
    csrc = os.path.join(SRC, "kernels", "selective_scan", "csrc", "selective_scan")
    build_dir = os.path.join(DST, "ext")
    os.makedirs(build_dir, exist_ok=True)
    nvcc = ["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
            "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__", "-U__CUDA_NO_BFLOAT162_OPERATORS__",
            "-U__CUDA_NO_BFLOAT162_CONVERSIONS__", "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math",
            "-lineinfo", "-gencode", "arch=compute_100,code=sm_100", "--threads", "4"]
    cpp_extension.load(
        name=EXT_NAME,
        sources=[os.path.join(csrc, "cus", f) for f in ("selective_scan.cpp", "selective_scan_core_fwd.cu", "selective_scan_core_bwd.cu")],
        extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=nvcc, extra_include_paths=[csrc],
        build_directory=build_dir, verbose=verbose, is_python_module=False)
    so = os.path.join(build_dir, EXT_NAME + ".so")
    assert os.path.exists(so), so
    # keep the .so only (objects and ninja files are build scratch and would travel to the GPU box for nothing)
    for f in os.listdir(build_dir):
        if not f.endswith(".so"):
            p = os.path.join(build_dir, f)
            shutil.rmtree(p) if os.path.isdir(p) else os.remove(p)
    return so


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-ext", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    if not os.path.isdir(SRC):
        print("install_ref: /root/reference/gm-unet not present (GPU box?) - nothing to do")
        return 0
    print("copied", copy_tree())
    if not args.no_ext:
        print("built", build_ext(args.verbose))
    return 0


if __name__ == "__main__":
    sys.exit(main())
