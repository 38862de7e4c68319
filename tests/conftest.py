import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """ctypes handle on the compiled, unmodified reference (oracle/_ref/libpbr_ref.so)."""
    from oracle import refharness as R
    if not R.available():
        pytest.skip("oracle/_ref/libpbr_ref.so not built (run oracle/build_ref.sh where /root/reference exists)")
    return R


@pytest.fixture(scope="session")
def api():
    from pysicalbasedraytracer_b200 import api as a
    return a


@pytest.fixture(scope="session")
def gpu_ctx(api):
    if api.device_count() == 0:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box (there is no CPU fallback)")
    return api.default_context(0)
