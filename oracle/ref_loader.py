"""Loads the UNMODIFIED reference classes from baseline/_ref/ (see install_ref.py) or, in the
authoring container, from /root/reference. TEST INFRASTRUCTURE / CPU-baseline arm only.

The reference's files are executed as they lie (importlib by file path): models/unet_model.py:65-146
(UNet), utils/losses.py:6-57 (WeightedCrossEntropyLoss), scripts/train.py:39-61 (center_crop_tensor,
init_weights). While scripts/train.py is executed, its `from models.unet_model import UNet` /
`from utils... import` lines are pointed at the reference's own modules, not at the drop-in shims of
this repository, so nothing of the product is on that path.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from typing import Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def _exec(name: str, path: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_root() -> Optional[str]:
    for base in CANDIDATES:
        if os.path.exists(os.path.join(base, "models", "unet_model.py")) and \
                os.path.exists(os.path.join(base, "utils", "losses.py")):
            return base
    return None


def load_reference() -> Optional[types.SimpleNamespace]:
    """Namespace(UNet, WeightedCrossEntropyLoss, init_weights, center_crop_tensor, root) or None."""
    base = reference_root()
    if base is None:
        return None
    sys.dont_write_bytecode = True          # /root/reference is read-only
    um = _exec("_reference_models_unet_model", os.path.join(base, "models", "unet_model.py"))
    lo = _exec("_reference_utils_losses", os.path.join(base, "utils", "losses.py"))
    ns = types.SimpleNamespace(UNet=um.UNet, WeightedCrossEntropyLoss=lo.WeightedCrossEntropyLoss,
                               init_weights=None, center_crop_tensor=None, root=base)
    train_py = os.path.join(base, "scripts", "train.py")
    if os.path.exists(train_py):
        names = ["models", "models.unet_model", "utils", "utils.losses", "utils.dataset",
                 "utils.augmentations"]
        saved = {k: sys.modules.get(k) for k in names}
        try:
            pm, pu = types.ModuleType("models"), types.ModuleType("utils")
            pm.__path__, pu.__path__ = [os.path.join(base, "models")], [os.path.join(base, "utils")]
            pm.unet_model, pu.losses = um, lo
            sys.modules.update({"models": pm, "models.unet_model": um, "utils": pu, "utils.losses": lo})
            sys.modules.pop("utils.dataset", None)
            sys.modules.pop("utils.augmentations", None)
            tr = _exec("_reference_scripts_train", train_py)     # body guarded by __main__ (:64)
            ns.init_weights, ns.center_crop_tensor = tr.init_weights, tr.center_crop_tensor
        except Exception:                                        # a missing optional dependency
            pass
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    return ns
