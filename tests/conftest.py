import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (B200); run with -m gpu")


def _cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def emu_lib():
    """CPU emulation build of the engine sources (tests/emu) -- test
    infrastructure for the kernels' logic; never used by the product."""
    from pypanadapter_b200 import _lib
    from tests.emu import build_emu
    return _lib.load_library(build_emu.build())


@pytest.fixture(scope="session")
def gpu_lib():
    if not _cuda_available():
        pytest.skip("no CUDA device")
    from pypanadapter_b200 import _lib
    return _lib.product_library()      # raises if the sm_100a library is missing


@pytest.fixture()
def emu_engine(emu_lib):
    from pypanadapter_b200.engine import ZoomPSD
    with ZoomPSD(0, lib=emu_lib) as e:
        yield e


@pytest.fixture()
def gpu_engine(gpu_lib):
    from pypanadapter_b200.engine import ZoomPSD
    with ZoomPSD(0, lib=gpu_lib) as e:
        yield e
