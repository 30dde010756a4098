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
    """Hash of everything the library is built from. Names are RELATIVE to the repository root: the
    stamp must stay valid when the tree is copied elsewhere (the GPU boxes run a snapshot under
    another path)."""
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
             if f.endswith((".cu", ".cuh", ".h"))] + [os.path.join(ROOT, "include", "unet_b200.h")]
    for path in files:
        with open(path, "rb") as fh:
            h.update(os.path.relpath(path, ROOT).replace(os.sep, "/").encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


class _BuildLock:
    """Inter-process lock around a (re)build: one rank of a multi-process launch compiles, the others
    wait and then load the finished library (the .so itself is replaced atomically)."""

    def __enter__(self):
        import fcntl

        os.makedirs(LIB_DIR, exist_ok=True)
        self.fh = open(os.path.join(LIB_DIR, "libunetb200.lock"), "w")
        fcntl.flock(self.fh, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl

        fcntl.flock(self.fh, fcntl.LOCK_UN)
        self.fh.close()
        return False


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
    with _BuildLock():
        digest = _source_hash()
        if not force and is_current():          # (re-checked under the lock: another rank may have built)
            LAST_ACTION = f"reused (hash {digest[:16]})"
            return LIB_PATH
        nvcc = _nvcc()
        tag = f"{os.getpid()}"

        def compile_one(src: str) -> str:
            obj = os.path.join(BUILD_DIR, src.replace(".cu", f".{tag}.o"))
            cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
            if verbose and r.stderr:
                print(r.stderr, file=sys.stderr)
            return obj

        with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
            objs = list(ex.map(compile_one, SOURCES))
        tmp = LIB_PATH + f".{tag}.tmp"
        cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        for o in objs:
            try:
                os.remove(o)
            except OSError:
                pass
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB_PATH)               # atomic: a concurrent loader never sees a partial file
        with open(STAMP_PATH + f".{tag}", "w") as fh:
            fh.write(digest)
        os.replace(STAMP_PATH + f".{tag}", STAMP_PATH)
        LAST_ACTION = f"rebuilt (hash {digest[:16]})"
        return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True), LAST_ACTION)
