"""Build bookkeeping of libunetb200 (no GPU): the source-hash stamp that `_lib.load()` consults must not
depend on where the tree lives (a GPU box runs a snapshot of the repository under another path — a
path-dependent stamp made every rank of a multi-process launch rebuild the library at once), and
concurrent loaders must serialise on the build lock."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_source_stamp_is_independent_of_the_checkout_path(tmp_path):
    from unet_segmentation_b200 import build

    if not os.path.exists(build.LIB_PATH):
        build.build()
    assert build.is_current()
    dst = tmp_path / "elsewhere" / "repo"
    shutil.copytree(os.path.join(ROOT, "unet_segmentation_b200"), dst / "unet_segmentation_b200",
                    ignore=shutil.ignore_patterns("__pycache__"))
    shutil.copytree(os.path.join(ROOT, "include"), dst / "include")
    r = subprocess.run([sys.executable, "-c",
                        "from unet_segmentation_b200 import build; print(build.is_current(), build._source_hash())"],
                       capture_output=True, text=True, cwd=str(dst), timeout=120)
    assert r.returncode == 0, r.stderr[-1000:]
    ok, digest = r.stdout.split()
    assert ok == "True" and digest == build._source_hash()


def test_concurrent_loaders_share_one_library(tmp_path):
    """Eight processes (one per rank of `torchrun --nproc-per-node 8`) load the library at the same
    time from a tree whose stamp is current: none rebuilds, all resolve the same symbols."""
    code = ("from unet_segmentation_b200 import _lib, build; lib = _lib.load(); "
            "print(build.LAST_ACTION or 'loaded', lib.ub_version())")
    procs = [subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True, cwd=ROOT) for _ in range(8)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-500:] for o in outs]
    assert all(o[0].split()[-1] == "100" and "rebuilt" not in o[0] for o in outs), [o[0] for o in outs]
