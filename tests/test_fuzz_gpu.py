"""Seeded random sweep of small shapes through the drop-in fwd/bwd against the C/f64 oracle: odd lengths (down to
L = 1), channel counts that do not fill a CTA tile, every state-count variant, optional arguments, all dtypes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GRADS = ["du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias"]


def rel_err(got, ref):
    got = got.detach().float().cpu().numpy().astype(np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def _cases():
    rng = np.random.RandomState(1234)
    out = []
    for i in range(48):
        G = int(rng.choice([1, 1, 2, 3, 4]))
        dpg = int(rng.choice([1, 2, 3, 5, 8, 16, 17, 33, 40]))
        N = int(rng.choice([1, 1, 2, 3, 4, 5, 8, 9, 16, 16, 17, 32, 33]))
        L = int(rng.choice([1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 200, 257]))
        nb = int(rng.choice([1, 2, 3]))
        dtype = [torch.float32, torch.float32, torch.bfloat16, torch.float16][i % 4]
        out.append((nb, G * dpg, L, N, G, bool(rng.randint(2)), bool(rng.randint(2)), bool(rng.randint(4) > 0), dtype, i))
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: "b%d_d%d_l%d_n%d_g%d_%d%d%d_%s_%d" % (
    c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], str(c[8]).split(".")[-1], c[9]))
def test_random_shapes(case):
    from ceigm_unet_b200.dropin import selective_scan_cuda_core as core
    from oracle import c_oracle
    from oracle.selective_scan_ref import make_inputs
    nb, nd, L, N, G, has_D, has_bias, sp, dtype, seed = case
    inp = make_inputs(nb, nd, L, N, groups=G, has_D=has_D, has_delta_bias=has_bias, dtype=dtype, seed=seed, device="cuda")
    out, x = core.fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], sp, 1)
    grads = core.bwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], inp["dout"], x, sp, 1)
    n = {k: (v.float().cpu().numpy() if v is not None else None) for k, v in inp.items()}
    ref_out, ref_last = c_oracle.scan_fwd(n["u"], n["delta"], n["A"], n["B"], n["C"], n["D"], n["delta_bias"], sp)
    ref_g = c_oracle.scan_bwd(n["u"], n["delta"], n["A"], n["B"], n["C"], n["D"], n["delta_bias"], n["dout"], sp)
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    assert rel_err(out, ref_out) < tol
    assert rel_err(x[:, :, -1, 1::2], ref_last) < tol
    for name, got in zip(GRADS, grads):
        if ref_g[name] is None:
            assert got is None, name
        else:
            assert rel_err(got, ref_g[name]) < tol, name
