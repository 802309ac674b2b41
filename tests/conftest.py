import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """One csv_ctx on cuda:0 for the whole session.  Without a usable B200 the tests that need it are SKIPPED (the library
    itself never falls back: csv_create fails with CSV_E_NODEVICE)."""
    from cutesv_b200 import _abi
    from cutesv_b200._lib import CuteSVError
    from cutesv_b200.engine import Engine
    try:
        e = Engine(0)
    except CuteSVError as err:
        if err.code == _abi.CSV_E_NODEVICE:
            pytest.skip("no usable sm_100 device: %s" % err)
        raise
    yield e
    e.close()
