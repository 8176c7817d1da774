"""FFN depthwise stack (csrc/ffn_dw.cu) next to the reference composition at the live stage shapes of a 224^2 batch-24 step:
python tools/bench_ffn.py [bf16|fp32]. CUDA events, median of 20 after 3 warm-ups; bytes = algorithmic passes x element size."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import ceigm_unet_b200 as pkg
from ceigm_unet_b200 import ops
from oracle import ffn_ref      # checker side: the reference composition timed next to ours

dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
dev = "cuda"


def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


print(f"# dtype {dt}; us per call; GB/s = algorithmic bytes (x in + y out, + aux) / time")
for name, C, H in (("s1 enc", 512, 56), ("s2 enc", 1024, 28), ("s3 enc", 1392, 14), ("s4 enc", 1792, 7), ("dec 56", 256, 56)):
    B, W = 24, H
    x = torch.randn(B, H * W, C, device=dev).to(dt)
    g = torch.randn(B, H * W, C, device=dev).to(dt)
    w3 = torch.randn(C, 1, 3, 3, device=dev); b3 = torch.randn(C, device=dev)
    gc = C // 8
    ms = [torch.randn(gc, 1, k, k, device=dev) for k in (3, 5, 7)]
    mb = [torch.randn(gc, device=dev) for _ in range(3)]
    segs = [(C - 3 * gc, 1, None, None), (C - 2 * gc, 3, ms[0], mb[0]), (C - gc, 5, ms[1], mb[1]), (C, 7, ms[2], mb[2])]
    es = x.element_size()
    nbytes = x.numel() * es
    t_gelu = timed(lambda: ops.dwnhwc_stencil(x, (H, W), [(C, 3, w3, b3)], epi=ops.EPI_GELU))
    t_ms = timed(lambda: ops.dwnhwc_stencil(x, (H, W), segs, epi=ops.EPI_RESIDUAL))
    t_dg = timed(lambda: ops.dwnhwc_stencil(x, (H, W), [(C, 3, w3, b3)], epi=ops.EPI_DGELU_MUL, aux=g))
    t_wg3 = timed(lambda: ops.dwnhwc_wgrad(x, g, (H, W), 0, C, 3))
    t_wg7 = timed(lambda: ops.dwnhwc_wgrad(x, g, (H, W), C - gc, C, 7))

    def ref_fwd():
        img = x.transpose(1, 2).reshape(B, C, H, W)
        y = F.gelu(F.conv2d(img, w3.to(dt), b3.to(dt), padding=1, groups=C))
        return y.flatten(2).transpose(1, 2).contiguous()
    t_ref = timed(ref_fwd)
    print(f"{name:7s} C={C:5d} {H}x{W}: dw3+GELU {t_gelu:7.1f} ({2*nbytes/t_gelu/1e3:6.0f} GB/s) | ref conv2d+GELU+transposes {t_ref:7.1f} | "
          f"multi-scale {t_ms:7.1f} ({2*nbytes/t_ms/1e3:6.0f}) | dGELU {t_dg:7.1f} ({3*nbytes/t_dg/1e3:6.0f}) | wgrad3 {t_wg3:7.1f} ({2*nbytes/t_wg3/1e3:6.0f}) | wgrad7(1/8) {t_wg7:7.1f}")
