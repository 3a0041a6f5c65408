"""In-tree build of libserenc.so (hand-written sm_100a CUDA behind a C ABI).

`python -m interspeech_ser_b200.build` or `build_library()`; nvcc cross-compiles without a GPU.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
REPO = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libserenc.so")
STAMP_PATH = os.path.join(PKG_DIR, ".libserenc.stamp")

SOURCES = ["serenc_api.cu"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libserenc cannot be built (there is no prebuilt or CPU fallback)")


def _source_digest() -> str:
    hsh = hashlib.sha256()
    # every file under csrc/ (a fixed header list once let edits to a new header go unbuilt)
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(REPO, "include", "serenc.h"))
    for f in files:
        with open(f, "rb") as fh:
            hsh.update(fh.read())
    hsh.update(b"AB_ARMS=" + os.environ.get("SERENC_AB_ARMS", "0").encode())   # compile flags are part of the identity
    return hsh.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into interspeech_ser_b200/libserenc.so for sm_100a. Returns the library path."""
    digest = _source_digest()

    def up_to_date() -> bool:
        if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
            return False
        with open(STAMP_PATH) as fh:
            return fh.read().strip() == digest

    if not force and up_to_date():
        return LIB_PATH
    # one builder at a time (N torchrun ranks import the package at once); the others wait and then find the stamp
    import fcntl
    lock = open(os.path.join(PKG_DIR, ".libserenc.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and up_to_date():
            return LIB_PATH
        return _compile(digest, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _compile(digest: str, verbose: bool) -> str:
    tmp_path = LIB_PATH + f".tmp{os.getpid()}"
    cmd = [
        _nvcc(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-std=c++17", "-lineinfo",
        "--shared", "-Xcompiler", "-fPIC",
        "-cudart", "shared",
        "-I", os.path.join(REPO, "include"),
        "-o", tmp_path,
    ]
    if os.environ.get("SERENC_AB_ARMS") == "1":   # development build: getenv kernel switches + the mma.sync attention arm
        cmd += ["-DSERENC_AB_ARMS"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libserenc.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    os.replace(tmp_path, LIB_PATH)   # atomic: a process that is loading the library sees the old or the new file, never half of one
    with open(STAMP_PATH, "w") as fh:
        fh.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
