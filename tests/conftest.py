import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port_oracle():
    from oracle.oracle_py import PortOracle
    return PortOracle()


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle.oracle_py import RefOracle
    if not RefOracle.available():
        pytest.skip("compiled reference (oracle/_ref/libnpref.so) not present on this box")
    return RefOracle()


@pytest.fixture(scope="session")
def engine():
    """One context on cuda:0 through the C ABI. No fallback: a missing library or device is an error."""
    from nanopolish_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()
