"""Install the UNMODIFIED reference hot path into baseline/_ref/ (git-ignored, travels to the GPU box).

The reference is a plain directory of Python scripts without packaging metadata, so the prescribed
`python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target
baseline/_ref /root/reference` stops with "Neither 'setup.py' nor 'pyproject.toml' found" (recorded in
DESIGN.md §6). This script does what that install would have done for the files on the hot path:
it copies them byte for byte from /root/reference (read-only; present in the authoring container
only) to baseline/_ref/, which is listed in .gitignore — no reference source enters the repository's
history — and not in .gpurunignore, so `bench.py --impl reference` and the `cpu_baseline` leg can
time the reference's own nn.Module and loss on the GPU box's host cores (`"kind": "reference"`).

Test infrastructure: nothing under unet_segmentation_b200/ imports baseline/_ref.
"""
from __future__ import annotations

import hashlib
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
# models/unet_model.py:5-146 (UNet), utils/losses.py:6-57 (WeightedCrossEntropyLoss),
# scripts/train.py:39-61 (center_crop_tensor, init_weights) and what importing train.py pulls in
FILES = ["models/unet_model.py", "utils/losses.py", "scripts/train.py", "utils/dataset.py",
         "utils/augmentations.py"]


def install(verbose: bool = True) -> bool:
    """Returns True when baseline/_ref holds the reference files (copied now or earlier)."""
    if not os.path.isdir(SRC):
        ok = all(os.path.exists(os.path.join(DST, f)) for f in FILES)
        if verbose:
            print(f"oracle/install_ref: {SRC} absent; baseline/_ref "
                  f"{'present (installed earlier)' if ok else 'absent'}")
        return ok
    manifest = []
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest.append(f"{hashlib.sha256(open(d, 'rb').read()).hexdigest()}  {f}")
    for pkg in ("models", "utils", "scripts"):
        init_src = os.path.join(SRC, pkg, "__init__.py")
        init_dst = os.path.join(DST, pkg, "__init__.py")
        if os.path.exists(init_src):
            shutil.copyfile(init_src, init_dst)
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    if verbose:
        print(f"oracle/install_ref: copied {len(FILES)} reference files to baseline/_ref/")
    return True


if __name__ == "__main__":
    install()
