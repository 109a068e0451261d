"""Shared pytest configuration: registers the `gpu` marker and makes the repo importable."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_available() -> bool:
    try:
        from hypredrive_b200 import hdk
        return hdk.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """Initialise the device runtime once; GPU tests fail loudly if the library is missing."""
    from hypredrive_b200 import hdk
    if hdk.device_count() <= 0:
        pytest.fail("test marked gpu but no CUDA device is visible")
    hdk.init()
    hdk.tune("amg_keep_debug", 1)       # tests compare S and the PMIS measures level by level
    return hdk
