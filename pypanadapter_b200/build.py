"""Build pypanadapter_b200/libzoomfft_b200.so in-tree with nvcc for sm_100a.

    python -m pypanadapter_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the
GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libzoomfft_b200.so")
SOURCES = [os.path.join(CSRC, "zfb_engine.cu")]
DEPS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + \
       [os.path.join(ROOT, "include", "zoomfft_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found; cannot build the sm_100a library")
    return nvcc


def source_hash() -> str:
    """sha256 over every source file's name and bytes and the compiler flags: the
    library carries it (zfb_source_hash), so "is this .so built from these
    sources" is a comparison, not a guess from mtimes."""
    import hashlib
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for d in DEPS:
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def built_hash() -> str | None:
    if not os.path.isfile(OUT):
        return None
    import ctypes
    try:
        lib = ctypes.CDLL(OUT, mode=ctypes.RTLD_LOCAL)
        fn = lib.zfb_source_hash
        fn.restype = ctypes.c_char_p
        return fn().decode()
    except (OSError, AttributeError):
        return None


def up_to_date() -> bool:
    return built_hash() == source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [find_nvcc(), *NVCC_FLAGS, '-DZFB_SOURCE_HASH="%s"' % source_hash(),
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, *SOURCES, "-o", OUT]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
