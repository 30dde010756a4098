"""Drop-in for the reference's ``utils/losses.py`` (``from utils.losses import
WeightedCrossEntropyLoss``, scripts/train.py:18)."""
from unet_segmentation_b200.loss import WeightedCrossEntropyLoss  # noqa: F401

__all__ = ["WeightedCrossEntropyLoss"]
