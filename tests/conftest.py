import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu")


def _has_gpu():
    if os.environ.get("ZKP_B200_FORCE_NO_GPU"):
        return False
    return os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")


@pytest.fixture(scope="session")
def native():
    """The product's native layer, initialised on cuda:0.  Fails loudly (no skip, no fallback)."""
    from interactive_zkp_study_b200 import native as nat
    nat.device_info()
    return nat
