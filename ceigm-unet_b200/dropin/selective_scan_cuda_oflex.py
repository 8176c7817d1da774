"""Drop-in for the reference extension `selective_scan_cuda_oflex`
(/root/reference/gm-unet/kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_oflex.cpp):
same as core plus `out_float`: the output (and `dout` in bwd) may be fp32 while u/delta/B/C are fp16/bf16.

    fwd(u, delta, A, B, C, D_, delta_bias_, delta_softplus, nrows, out_float) -> [out, x]
    bwd(u, delta, A, B, C, D_, delta_bias_, dout, x_, delta_softplus, nrows) -> [du, ddelta, dA, dB, dC, dD, ddelta_bias]
"""
import torch

from .. import ops


def fwd(u, delta, A, B, C, D_=None, delta_bias_=None, delta_softplus=False, nrows=1, out_float=True):
    out, x = ops.ScanProblem(u, delta, A, B, C, D_, delta_bias_, delta_softplus, out_float=bool(out_float)).forward(True)
    return [out, x]


def bwd(u, delta, A, B, C, D_, delta_bias_, dout, x_=None, delta_softplus=False, nrows=1):
    out_float = dout.dtype == torch.float32 and u.dtype != torch.float32     # the reference infers it the same way
    return ops.ScanProblem(u, delta, A, B, C, D_, delta_bias_, delta_softplus, out_float=out_float).backward(dout, x_)
