import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port_lib():
    from oracle import okv
    okv.build(ref=os.path.isdir("/root/reference"))
    return okv.load_port()


@pytest.fixture(scope="session")
def ref_lib(port_lib):
    from oracle import okv
    if not okv.ref_available():
        pytest.skip("oracle/_ref (the reference's own RocksDB binary) not built here")
    return okv.load_ref()


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """librsp_b200.so is built in-tree (nvcc, sm_100a); rebuild only when sources are newer.  No GPU needed."""
    from rocksplicator_b200 import build
    build.build()
