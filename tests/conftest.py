import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracles():
    """Compile the CPU checkers once (seconds).  The reference half is only (re)built where
    /root/reference exists; on the GPU box the prebuilt oracle/_ref/ travels with the snapshot."""
    from oracle import pyoracle as orc
    if not orc.available("port") or (os.path.isdir("/root/reference") and not orc.available("ref")):
        orc.build()
    yield
