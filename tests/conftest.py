import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import orc as _orc
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def rast_factory():
    """Creates swr::Rasterizer mirrors on cuda:0; fails loudly when the CUDA library is missing."""
    from glimpsw_b200 import api

    created = []

    def make(**kw):
        r = api.Rasterizer(0, **kw)
        created.append(r)
        return r

    yield make
    for r in created:
        r.destroy()
