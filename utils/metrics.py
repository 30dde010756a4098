"""Drop-in for the hot-path part of the reference's ``utils/metrics.py``:
``get_instance_masks`` (utils/metrics.py:42-72, called by scripts/predict.py:96-98) runs on the GPU.
The evaluation-only helpers of the reference's module (``calculate_iou``,
``calculate_rand_index_and_error`` — out of scope) are forwarded on first use to the
reference's own file when it is reachable through ``sys.path``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

from unet_segmentation_b200.postprocess import get_instance_masks  # noqa: F401

__all__ = ["get_instance_masks"]
_reference_module = None


def _load_reference_metrics():
    global _reference_module
    if _reference_module is None:
        here = _os.path.abspath(__file__)
        for entry in _sys.path:
            cand = _os.path.abspath(_os.path.join(entry or ".", "utils", "metrics.py"))
            if cand != here and _os.path.isfile(cand):
                spec = _ilu.spec_from_file_location("_reference_utils_metrics", cand)
                mod = _ilu.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _reference_module = mod
                break
    return _reference_module


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    mod = _load_reference_metrics()
    if mod is None or not hasattr(mod, name):
        raise AttributeError(f"module 'utils.metrics' has no attribute {name!r} (only "
                             "get_instance_masks is provided by the B200 drop-in; the reference's "
                             "utils/metrics.py was not found on sys.path)")
    return getattr(mod, name)
