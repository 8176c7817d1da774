"""Minimal stand-in for `timm` (not installed in this image; no network). Only what the reference GM-UNet imports:
model/gm/groupmamba.py:6-8, model/gm/ss2d.py:14, model/gm/custom_mlp.py:3, model/best_decoder.py:7-8,
model/vmamba/vmamba.py:11. Behaviour follows timm's public definitions of these helpers."""
from . import models  # noqa: F401
