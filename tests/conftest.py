import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "volume-renderer_b200", "python"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle and the product libraries exist (built in-tree; on the GPU box the
    prebuilt .so files travel with the snapshot)."""
    from oracle import orc
    orc.build()
    import volren_b200
    if not (os.path.exists(volren_b200.LIB_PATH) and os.path.exists(volren_b200.HOST_LIB_PATH)):
        volren_b200.build()
    yield


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
