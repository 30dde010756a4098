"""The zero-edit drop-in promise of INTEGRATION.md §1: with this repository in front of the
reference on ``sys.path`` the reference's scripts import the B200 ``UNet`` /
``WeightedCrossEntropyLoss`` / ``get_instance_masks``, while every module this repository does not
replace (``utils.dataset``, ``utils.augmentations`` …) still comes from the reference tree."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _run(code, pythonpath, cwd):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath), PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, cwd=cwd,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_shim_packages_fall_through_to_a_reference_shaped_tree(tmp_path):
    """A stand-in tree laid out like the reference (``utils/`` and ``models/`` WITHOUT __init__.py)."""
    (tmp_path / "utils").mkdir()
    (tmp_path / "models").mkdir()
    (tmp_path / "utils" / "dataset.py").write_text("MARK = 'reference dataset'\n")
    (tmp_path / "utils" / "losses.py").write_text("MARK = 'reference losses'\n")
    (tmp_path / "utils" / "metrics.py").write_text("def calculate_iou(a, b):\n    return 'reference iou'\n")
    (tmp_path / "models" / "unet_model.py").write_text("MARK = 'reference model'\n")
    (tmp_path / "models" / "other.py").write_text("MARK = 'reference other'\n")
    out = _run("""
        import sys
        sys.path.insert(0, sys.argv[0] if False else %r)   # what scripts/train.py:11-14 does
        from utils.dataset import MARK as d
        from models.other import MARK as o
        from models.unet_model import UNet
        from utils.losses import WeightedCrossEntropyLoss
        import utils.metrics as m
        print(d, '|', o, '|', UNet.__module__, '|', WeightedCrossEntropyLoss.__module__, '|',
              m.get_instance_masks.__module__, '|', m.calculate_iou(0, 0))
        """ % str(tmp_path), [ROOT, str(tmp_path)], str(tmp_path))
    assert out.strip() == ("reference dataset | reference other | unet_segmentation_b200.unet | "
                           "unet_segmentation_b200.loss | unet_segmentation_b200.postprocess | "
                           "reference iou")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_reference_train_script_binds_to_the_drop_in():
    """Import the reference's unmodified scripts/train.py (its body is guarded by __main__)."""
    out = _run("""
        import importlib.util
        spec = importlib.util.spec_from_file_location('ref_train', %r)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        print(m.UNet.__module__, m.WeightedCrossEntropyLoss.__module__, m.HeLaDataset.__module__,
              m.HeLaDataset.__init__.__code__.co_filename)
        """ % os.path.join(REF, "scripts", "train.py"), [ROOT], "/tmp")
    mods = out.split()
    assert mods[:3] == ["unet_segmentation_b200.unet", "unet_segmentation_b200.loss", "utils.dataset"]
    assert mods[3] == os.path.join(REF, "utils", "dataset.py")
