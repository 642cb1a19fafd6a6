"""pytest configuration: `gpu` marker, shared fixtures.

`-m "not gpu"` runs the oracle against the golden vectors / the verbatim-compiled reference, the
host-side logic and the C-ABI symbol check (no compute calls).  `-m gpu` runs the parity tests
proper through the C ABI on a B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_available():
    # A box with an NVIDIA device node runs the GPU tests unconditionally: a missing or broken
    # libcavb200.so must then FAIL them, never skip them.
    if os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0"):
        return True
    try:
        from cav_hoomd_b200 import capi
        return capi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests must not silently pass on a box without a GPU: skip them loudly here (CPU box);
    # on the GPU box a missing library is a hard failure inside the test itself.
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def coracle():
    from oracle import oracle as O
    return O.COracle()


@pytest.fixture(scope="session")
def reforacle():
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return O.RefOracle()


@pytest.fixture(scope="session")
def handle():
    from cav_hoomd_b200 import capi
    h = capi.Handle(0)
    yield h
    h.close()
