"""Host-side checks that need no GPU: the C-ABI library loads and exports every
symbol include/zoomfft_b200.h declares, its host-only entry points agree with
the oracle, and there is no CPU fallback."""
import ctypes as C
import os
import re
import shutil

import numpy as np
import pytest

from oracle import zoompsd_oracle as zo
from pypanadapter_b200 import _lib, build, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "zoomfft_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zfb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def product_lib():
    if not os.path.isfile(_lib.LIB_PATH):
        if shutil.which("nvcc") is None and not os.path.isfile("/usr/local/cuda/bin/nvcc"):
            pytest.skip("no nvcc and no prebuilt library")
        build.build()
    return _lib.product_library()


def test_binding_covers_header():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_symbol(product_lib):
    for name in declared_symbols():
        assert hasattr(product_lib, name), name
    assert product_lib.zfb_abi_version() == _lib.ABI_VERSION
    assert product_lib.zfb_build_kind() == b"sm_100a"


def test_struct_layout_matches_header():
    # double, 9 x int32, (pad), 2 x double, pointer  -> 72 bytes on LP64
    assert C.sizeof(_lib.ZfbConfig) == 72
    assert _lib.ZfbConfig.f_demod.offset == 48
    assert _lib.ZfbConfig.window.offset == 64


def test_sos_matches_scipy(product_lib):
    assert np.abs(engine.decim_sos(product_lib) - zo.decim_sos()).max() < 1e-13


@pytest.mark.parametrize("n,N,R", [(239616, 2048, 8), (319488, 4096, 16), (100003, 2048, 8),
                                   (33333, 1024, 4), (8192, 2048, 8), (4096, 1024, 4),
                                   (1 << 20, 65536, 1), (2048, 32, 2), (28, 32, 2), (57, 64, 3)])
def test_plan_geometry_matches_scipy_lengths(product_lib, n, N, R):
    g = engine.plan_geometry(n, N, R, product_lib)
    L = n
    for _ in range(int(np.log2(R))):
        L = len(np.zeros(L)[::2])
    nperseg, hop, nseg = zo.welch_plan(L, N)
    assert (g["ndec"], g["nperseg"], g["hop"], g["nseg"]) == (L, nperseg, hop, nseg)
    assert g["nstages"] == int(np.log2(R))


def test_plan_geometry_rejects_short(product_lib):
    with pytest.raises(ValueError):
        engine.plan_geometry(27, 32, 2, product_lib)
    with pytest.raises(ValueError):
        engine.plan_geometry(54, 32, 4, product_lib)      # second stage sees 27


def test_no_cpu_fallback(product_lib):
    """Without a CUDA device engine creation must fail loudly."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("CUDA device present")
    except ImportError:
        pass
    with pytest.raises(engine.ZoomFFTError) as ei:
        engine.ZoomPSD(0, lib=product_lib)
    assert ei.value.code == _lib.ZFB_ENODEV
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pypanadapter_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libzoomfft_emu" not in src, f


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 on its own."""
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
