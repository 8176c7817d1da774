import sys, statistics
sys.path.insert(0, "/root/repo")
import torch
from ceigm_unet_b200 import ops
torch.manual_seed(0)
b, dt, L, N, G = 24, 768, 3136, 16, 4
A = -0.5 * torch.rand(dt, N, device="cuda")
D = torch.randn(dt, device="cuda"); bias = 0.5 * torch.rand(dt, device="cuda")
def mk(dtype):
    g = torch.Generator(device="cuda").manual_seed(0)
    return dict(u=torch.randn(b, dt, L, device="cuda", generator=g).to(dtype), dl=(0.5 * torch.rand(b, dt, L, device="cuda", generator=g)).to(dtype),
                B=torch.randn(b, G, N, L, device="cuda", generator=g).to(dtype), C=torch.randn(b, G, N, L, device="cuda", generator=g).to(dtype),
                dout=torch.randn(b, dt, L, device="cuda", generator=g).to(dtype))
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
for dtype in (torch.float32, torch.bfloat16, torch.float16):
    i = mk(dtype)
    pr = ops.ScanProblem(i["u"], i["dl"], A, i["B"], i["C"], D, bias, True)
    out, x = pr.forward(True)
    f = t(lambda: pr.forward(True))
    bw = t(lambda: pr.backward(i["dout"], x))
    print(dtype, "fwd %.3f ms  bwd %.3f ms" % (f, bw))
