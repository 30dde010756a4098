"""Build libunetb200.so (sm_100a only) in-tree with nvcc.

Usage: python -m unet_segmentation_b200.build [--force]
The shared library is written to unet_segmentation_b200/lib/libunetb200.so; object files go to
build/ (git-ignored). nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libunetb200.so")
BUILD_DIR = os.path.join(ROOT, "build", "libunetb200")
SOURCES = ["igemm.cu", "kernels.cu", "capi_ops.cu", "plan.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libunetb200 cannot be built (there is no CPU fallback)")


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(ROOT, "include", "unet_b200.h")]
    for f in files:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(path, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


STAMP_PATH = os.path.join(LIB_DIR, "libunetb200.hash")
LAST_ACTION = ""     # "rebuilt (hash ...)" / "reused (hash ...)": what the last build() call did


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def is_current() -> bool:
    """Does the in-tree library carry the hash stamp of the sources as they are now?"""
    try:
        return os.path.exists(LIB_PATH) and open(STAMP_PATH).read().strip() == _source_hash()
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    global LAST_ACTION
    os.makedirs(BUILD_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = STAMP_PATH
    digest = _source_hash()
    if not force and is_current():
        LAST_ACTION = f"reused (hash {digest[:16]})"
        return LIB_PATH
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr, file=sys.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    LAST_ACTION = f"rebuilt (hash {digest[:16]})"
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True), LAST_ACTION)
