import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def port():
    """The CPU oracle (oracle/port_oracle.py); builds the C port on first use."""
    from oracle import port_oracle
    port_oracle.lib()
    return port_oracle


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref); skipped where it was not built."""
    from oracle import ref_oracle
    if not ref_oracle.available():
        pytest.skip("oracle/_ref/liboracle_ref.so not built here")
    ref_oracle.lib()
    return ref_oracle


@pytest.fixture(scope="session")
def gpu():
    """The product: annongpu_b200 over libangpu.so on cuda:0."""
    import annongpu_b200 as A
    A.setDevice(0)
    return A
