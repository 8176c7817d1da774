"""fwd+bwd time of the live GM-UNet block `GroupMambaLayer` (4 single-direction SS2Ds, d_state = 1; groupmamba.py:85-159)
on the four stage shapes of a 224^2 batch-24 step: eager launches vs whole-layer CUDA graphs (forward graph + backward
graph captured by torch.cuda.make_graphed_callables around our autograd Functions). The live regime is launch-bound
(SURVEY.md §8-f2): the graph removes the ~120 host launches per layer call from the critical path.
    python tools/bench_gml.py [batch]"""
import copy, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import ceigm_unet_b200 as P

Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
STAGES = [(64, 56, 5), (128, 28, 6), (348, 14, 12), (448, 7, 3)]      # (channels C = 4 d_model, H = W, layers per forward)
torch.set_float32_matmul_precision("medium")      # what the reference trains with (train_synapse.py:21)
torch.manual_seed(0)


class Wrap(torch.nn.Module):
    def __init__(self, layer, H):
        super().__init__()
        self.layer, self.H = layer, H

    def forward(self, x):
        return self.layer(x, self.H, self.H)


def time_step(fn, x, gy, params, iters=20):
    def step():
        y = fn(x)
        y.backward(gy)
        x.grad = None
        for p_ in params:
            p_.grad = None
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


tot_e = tot_g = 0.0
for C, H, count in STAGES:
    m = Wrap(P.GroupMambaLayer(C, C), H).cuda()
    mg = copy.deepcopy(m)          # same weights; captured before any eager backward touches its parameters (see graphed())
    x = torch.randn(Bn, H * H, C, device="cuda", requires_grad=True)
    gy = torch.randn(Bn, H * H, C, device="cuda")
    g = P.graphed(mg, (x.detach().clone().requires_grad_(True),))
    y_g = g(x); y_g.backward(gy); gx_g = x.grad.clone(); x.grad = None
    t_g = time_step(g, x, gy, list(mg.parameters()))
    params = list(m.parameters())
    P.launch_count(reset=True)
    y_ref = m(x); y_ref.backward(gy); gx_ref = x.grad.clone(); x.grad = None
    launches = P.launch_count()
    err = max(((y_g - y_ref).abs().max() / y_ref.abs().max()).item(), ((gx_g - gx_ref).abs().max() / gx_ref.abs().max()).item())
    t_e = time_step(m, x, gy, params)
    tot_e += count * t_e; tot_g += count * t_g
    print(f"GroupMambaLayer C={C} {H}x{H} B={Bn}: eager {t_e:.3f} ms | graphed {t_g:.3f} ms | x{t_e / t_g:.2f} | "
          f"our kernel launches per fwd+bwd {launches} | max rel diff graphed vs eager {err:.1e}")
print(f"all {sum(c for _, _, c in STAGES)} GroupMambaLayer calls of one 224^2 step (B={Bn}): eager {tot_e:.2f} ms | graphed {tot_g:.2f} ms "
      f"-> {Bn / (tot_g * 1e-3):.0f} slices/s (SSM branch only)")
