"""Minimal driver for ncu captures of the FFN depthwise-stack kernels (csrc/ffn_dw.cu): one custom_ffn-sized call of every
kernel at the stage-1 encoder shape of a 224^2 batch-24 step (hidden 512 @ 56^2, bf16 by default).
    python tools/prof_ffn.py [bf16|fp32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ceigm_unet_b200 import ops  # noqa: E402

dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
B, H, W, C = 24, 56, 56, 512
x = torch.randn(B, H * W, C, device="cuda").to(dt)
g = torch.randn(B, H * W, C, device="cuda").to(dt)
w3 = torch.randn(C, 1, 3, 3, device="cuda"); b3 = torch.randn(C, device="cuda")
gc = C // 8
ms = [torch.randn(gc, 1, k, k, device="cuda") for k in (3, 5, 7)]
mb = [torch.randn(gc, device="cuda") for _ in range(3)]
segs = [(C - 3 * gc, 1, None, None), (C - 2 * gc, 3, ms[0], mb[0]), (C - gc, 5, ms[1], mb[1]), (C, 7, ms[2], mb[2])]
for _ in range(2):
    ops.dwnhwc_stencil(x, (H, W), [(C, 3, w3, b3)], epi=ops.EPI_GELU)
    ops.dwnhwc_stencil(x, (H, W), segs, epi=ops.EPI_RESIDUAL)
    ops.dwnhwc_stencil(x, (H, W), [(C, 3, w3, b3)], epi=ops.EPI_DGELU_MUL, aux=g)
    ops.dwnhwc_wgrad(x, g, (H, W), 0, C, 3)
    ops.dwnhwc_wgrad(x, g, (H, W), C - gc, C, 7)
torch.cuda.synchronize()
print("ok")
