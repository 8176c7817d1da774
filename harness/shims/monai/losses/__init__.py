from . import focal_loss  # noqa: F401
