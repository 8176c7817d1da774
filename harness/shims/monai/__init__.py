"""Stand-in for `monai` (absent): the reference's loss.py:5 imports FocalLoss at module level; only DiceFocalLoss uses it
and no GM-UNet training script selects that loss (train_synapse.py:90-93 uses DiceCELoss)."""
from . import losses  # noqa: F401
