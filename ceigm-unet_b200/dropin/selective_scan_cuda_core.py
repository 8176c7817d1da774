"""Drop-in for the reference extension `selective_scan_cuda_core`
(/root/reference/gm-unet/kernels/selective_scan/csrc/selective_scan/cus/selective_scan.cpp:157-164, 241-250, 351-354).

    fwd(u, delta, A, B, C, D_, delta_bias_, delta_softplus, nrows) -> [out, x]
    bwd(u, delta, A, B, C, D_, delta_bias_, dout, x_, delta_softplus, nrows) -> [du, ddelta, dA, dB, dC, dD, ddelta_bias]

`nrows` is accepted and ignored, as in the reference (it always instantiates <1, ...>, selective_scan.cpp:235).
`x` keeps the reference convention x[:, :, -1, 1::2] == final state; the chunk checkpoints used by `bwd` sit in
front of it in the same storage, and `bwd` recomputes them when handed a foreign or missing `x_`.
"""
from .. import ops


def fwd(u, delta, A, B, C, D_=None, delta_bias_=None, delta_softplus=False, nrows=1):
    out, x = ops.ScanProblem(u, delta, A, B, C, D_, delta_bias_, delta_softplus, out_float=False).forward(True)
    return [out, x]


def bwd(u, delta, A, B, C, D_, delta_bias_, dout, x_=None, delta_softplus=False, nrows=1):
    return ops.ScanProblem(u, delta, A, B, C, D_, delta_bias_, delta_softplus, out_float=False).backward(dout, x_)
