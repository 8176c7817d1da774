"""cProfile of the host side of one eager GroupMambaLayer fwd+bwd (stage-4 shape: the GPU work is negligible)."""
import cProfile, pstats, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import ceigm_unet_b200 as P
torch.manual_seed(0)
m = P.GroupMambaLayer(448, 448).cuda()
x = torch.randn(24, 49, 448, device="cuda", requires_grad=True)
gy = torch.randn(24, 49, 448, device="cuda")
def step():
    y = m(x, 7, 7); y.backward(gy)
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
