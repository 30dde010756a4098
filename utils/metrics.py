"""Drop-in for the hot-path part of the reference's ``utils/metrics.py``:
``get_instance_masks`` (utils/metrics.py:42-72, called by scripts/predict.py:96-98)."""
from unet_segmentation_b200.postprocess import get_instance_masks  # noqa: F401

__all__ = ["get_instance_masks"]
