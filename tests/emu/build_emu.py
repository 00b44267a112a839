"""TEST INFRASTRUCTURE ONLY: build the CPU emulation of the engine sources.

Compiles pypanadapter_b200/csrc/zfb_engine.cu (the same file nvcc builds for
sm_100a) with g++ -DZFB_EMULATE against tests/emu/cuda_emu.h into
tests/emu/_build/libzoomfft_emu.so.  Only tests/test_emu_*.py load it.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pypanadapter_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libzoomfft_emu.so")


def build(force: bool = False) -> str:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "cuda_emu.cpp"),
            os.path.join(ROOT, "include", "zoomfft_b200.h")]
    if (not force and os.path.isfile(OUT)
            and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps)):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-mfma", "-fPIC", "-shared", "-DZFB_EMULATE",
           "-I", HERE, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-x", "c++", os.path.join(CSRC, "zfb_engine.cu"),
           "-x", "c++", os.path.join(HERE, "cuda_emu.cpp"), "-o", OUT, "-lpthread"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
