import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


def _cuda_device_count() -> int:
    try:
        import ctypes
        lib = ctypes.CDLL("libcudart.so.12")
    except OSError:
        try:
            import torch
            return torch.cuda.device_count()
        except Exception:
            return 0
    n = ctypes.c_int(0)
    rc = lib.cudaGetDeviceCount(ctypes.byref(n))
    return n.value if rc == 0 else 0


@pytest.fixture(scope="session")
def gpu_required():
    """GPU tests must not silently skip on a GPU box: they fail if the extension is missing there."""
    from pimd_b_b200 import _cabi
    _cabi.load()   # raises if libpimdb200.so has not been built
    return True
