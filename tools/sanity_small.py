"""Small end-to-end exercise of every kernel (for compute-sanitizer): tiny shapes, all code paths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import ceigm_unet_b200 as P
from ceigm_unet_b200.dropin import selective_scan_cuda_core as core
torch.manual_seed(0)
dev = "cuda:0"
def case(b, d, L, n, g, dtype=torch.float32):
    A = -0.5 * torch.rand(d, n, device=dev)
    Bm, Cm = torch.randn(b, g, n, L, device=dev).to(dtype), torch.randn(b, g, n, L, device=dev).to(dtype)
    u, dl = torch.randn(b, d, L, device=dev).to(dtype), (0.5 * torch.rand(b, d, L, device=dev)).to(dtype)
    D_, bias = torch.randn(d, device=dev), 0.5 * torch.rand(d, device=dev)
    out, x = core.fwd(u, dl, A, Bm, Cm, D_, bias, True, 1)
    core.bwd(u, dl, A, Bm, Cm, D_, bias, torch.randn_like(out), x, True, 1)
    core.bwd(u, dl, A, Bm, Cm, D_, bias, torch.randn_like(out), None, True, 1)
for args in [(2, 40, 100, 16, 2), (1, 24, 64, 16, 1), (1, 72, 160, 12, 2), (2, 12, 77, 3, 1), (2, 16, 196, 1, 1), (1, 8, 49, 1, 1), (1, 6, 40, 48, 1),
             (2, 40, 128, 16, 2, torch.bfloat16), (2, 16, 196, 1, 1, torch.bfloat16), (1, 70, 3136, 1, 1)]:
    case(*args)
for k, N, dm in [(4, 16, 8), (1, 1, 8)]:
    m = P.SS2D(d_model=dm, d_state=N, ssm_ratio=2.0 if k == 4 else 1, k_group=k).cuda()
    x = torch.randn(2, 6, 5, dm, device=dev, requires_grad=True)
    y = m(x) if k == 4 else m(x, CrossScan=P.CrossScan_4, CrossMerge=P.CrossMerge_4)
    y.sum().backward()
layer = P.GroupMambaLayer(32, 32).cuda()
layer(torch.randn(2, 36, 32, device=dev), 6, 6).sum().backward()
xs = P.CrossScan.apply(torch.randn(2, 3, 5, 6, device=dev)); P.CrossMerge.apply(xs.view(2, 4, 3, 5, 6))
torch.cuda.synchronize()
print("sanity ok")
