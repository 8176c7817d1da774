import sys
sys.path.insert(0, "/root/repo")
import torch, ctypes
from ceigm_unet_b200 import ops, _lib
Bn, K, D, H, W = 24, 4, 192, 56, 56
L = H * W
ys = torch.randn(Bn, K, D, L, device="cuda")
lnw, lnb = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
for dtype, code in ((torch.bfloat16, _lib.SS2D_BF16), (torch.float32, _lib.SS2D_F32)):
    z = torch.randn(Bn, L, D, device="cuda").to(dtype)
    out = torch.empty(Bn, L, D, device="cuda", dtype=dtype)
    stats = torch.empty(Bn, L, 2, device="cuda")
    for _ in range(2):
        rc = _lib.lib().ss2d_gate_proj_fwd(ctypes.c_void_p(ys.data_ptr()), K, ctypes.c_uint32(0b1010), ctypes.c_void_p(lnw.data_ptr()),
                                           ctypes.c_void_p(lnb.data_ptr()), ctypes.c_float(1e-5), ctypes.c_void_p(z.data_ptr()), D, 1, None, 0, None,
                                           None, 0, ctypes.c_void_p(out.data_ptr()), D, ctypes.c_void_p(stats.data_ptr()), Bn, D, L, H, W, 0,
                                           code, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
torch.cuda.synchronize()
print("ok")
