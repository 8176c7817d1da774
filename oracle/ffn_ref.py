"""ORACLE (test infrastructure only — imported by tests/, never by the product): plain-PyTorch restatement of the
reference's feed-forward blocks next to the SS2D branch, written against a state_dict `p`.

  pvt2_ffn    PVT2FFN.forward              /root/reference/gm-unet/model/gm/groupmamba.py:77-82 (+ DWConv.forward :450-455)
  custom_ffn  custom_ffn.forward           /root/reference/gm-unet/model/gm/custom_mlp.py:362-368
              InceptionDWConv2d_MultiScale /root/reference/gm-unet/model/gm/custom_mlp.py:323-336

Pinned: tests/golden/ffn_pvt2.npz and ffn_custom.npz hold inputs, outputs and every gradient recorded from the unmodified
reference modules (tests/golden/make_golden.py: gen_ffn); tests/test_oracle_golden.py checks these functions against them.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _dwconv(x, p, H, W):
    """DWConv.forward (groupmamba.py:450-455): (B, L, C) -> NCHW -> depthwise 3 x 3 + bias -> (B, L, C)."""
    Bn, L, C = x.shape
    img = x.transpose(1, 2).reshape(Bn, C, H, W)
    img = F.conv2d(img, p["dwconv.dwconv.weight"], p["dwconv.dwconv.bias"], stride=1, padding=1, groups=C)
    return img.flatten(2).transpose(1, 2)


def pvt2_ffn(x, p, H, W):
    """x (B, L, C) -> (B, L, C). groupmamba.py:77-82."""
    h = F.linear(x, p["fc1.weight"], p["fc1.bias"])          # :78
    h = _dwconv(h, p, H, W)                                  # :79
    h = F.gelu(h)                                            # :80 (nn.GELU(): exact erf)
    return F.linear(h, p["fc2.weight"], p["fc2.bias"])       # :81


def multi_scale(x, p, H, W, prefix="custom."):
    """InceptionDWConv2d_MultiScale.forward (custom_mlp.py:323-336)."""
    Bn, L, C = x.shape
    gc = p[prefix + "dwconv_3x3.weight"].shape[0]
    img = x.transpose(1, 2).reshape(Bn, C, H, W)
    x_id, x3, x5, x7 = torch.split(img, (C - 3 * gc, gc, gc, gc), dim=1)                                   # :330
    o3 = F.conv2d(x3, p[prefix + "dwconv_3x3.weight"], p[prefix + "dwconv_3x3.bias"], padding=1, groups=gc)  # :333
    o5 = F.conv2d(x5, p[prefix + "dwconv_5x5.weight"], p[prefix + "dwconv_5x5.bias"], padding=2, groups=gc)  # :334
    o7 = F.conv2d(x7, p[prefix + "dwconv_7x7.weight"], p[prefix + "dwconv_7x7.bias"], padding=3, groups=gc)  # :335
    return (img + torch.cat((x_id, o3, o5, o7), dim=1)).flatten(2).transpose(1, 2)                         # :336


def custom_ffn(x, p, H, W):
    """x (B, L, C) -> (B, L, C). custom_mlp.py:362-368."""
    h = F.linear(x, p["fc1.weight"], p["fc1.bias"])          # :363
    h = _dwconv(h, p, H, W)                                  # :364
    h = F.gelu(h)                                            # :365
    h = multi_scale(h, p, H, W)                              # :366
    return F.linear(h, p["fc2.weight"], p["fc2.bias"])       # :367
