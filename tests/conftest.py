import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import pyoracle
    pyoracle.build()
    if not pyoracle.Ref.available():
        pytest.skip("oracle/_ref/libatlas_ref.so not built (needs /root/reference)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def ctx():
    import __graft_entry__ as g
    g.build()
    from atlas_engine_b200 import capi
    c = capi.Context(0)   # raises loudly without a CUDA device: there is no CPU fallback
    yield c
    c.close()
