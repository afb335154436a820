import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes mirror over librl_b200.so), built if stale."""
    entry.build_library()
    entry.build_host_library()
    return entry.load_package()


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (checker)."""
    entry.build_oracle()
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def gpu(pkg):
    if pkg.device_count() < 1:
        pytest.skip("no CUDA device")
    return pkg
