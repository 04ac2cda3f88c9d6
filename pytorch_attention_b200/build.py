"""In-tree build of libpa_b200.so with nvcc for sm_100a (no JIT cache, no torch extension machinery).

The built library stays next to the sources (pytorch_attention_b200/lib/) so that it travels with the
repository snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libpa_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build libpa_b200.so")
    return cand


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(INCLUDE_DIR, "*.h")))


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _deps())


def build_lib(force=False, verbose=False):
    """Compile csrc/*.cu into lib/libpa_b200.so.  Returns the library path.

    Safe with several processes (N torchrun ranks finding a stale library): the build runs under an exclusive file lock,
    nvcc writes to a temporary name and the result is moved into place atomically, so nobody can dlopen a half-written
    file; a rank that waited for the lock re-checks staleness and skips the compile."""
    if not force and not is_stale():
        return LIB_PATH
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB_PATH                  # another process built it while we waited
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + sources()
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def header_version():
    """PA_VERSION of include/pa_b200.h (what a freshly built library reports from pa_version())."""
    import re
    with open(os.path.join(INCLUDE_DIR, "pa_b200.h")) as f:
        m = re.search(r"#define\s+PA_VERSION\s+(\d+)", f.read())
    return int(m.group(1)) if m else None


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
