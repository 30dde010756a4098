"""Drop-in for the reference's ``models/unet_model.py``: ``from models.unet_model import UNet`` in
scripts/train.py:17 / scripts/predict.py:21 resolves here when this repository precedes the
reference on ``sys.path``. The implementation lives in ``unet_segmentation_b200.unet``."""
from unet_segmentation_b200.unet import DoubleConv, Down, OutConv, UNet, Up  # noqa: F401

__all__ = ["UNet", "DoubleConv", "Down", "Up", "OutConv"]
