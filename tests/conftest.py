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


def pytest_sessionfinish(session, exitstatus):
    """ZKP_B200_GUARD=1 python -m pytest tests -m gpu: every device block of the library carries canary zones;
    a single overwritten canary byte fails the run (compute-sanitizer is closed on the B200 pool)."""
    if not os.environ.get("ZKP_B200_GUARD") or not _has_gpu():
        return
    from interactive_zkp_study_b200 import _lib, native as nat
    if not _lib._initialised:
        return
    live, bad = nat.debug_check_guards()
    print("\n[zkp guard] %d live device blocks checked, %d damaged" % (live, bad))
    if bad:
        session.exitstatus = 1
