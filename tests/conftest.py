import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the ARIMA oracle evaluates likelihoods at degenerate parameters (overflow -> -1e300 by design)
    config.addinivalue_line("filterwarnings", "ignore::RuntimeWarning")
    config.addinivalue_line("filterwarnings", "ignore::UserWarning:scipy")


@pytest.fixture(scope="session")
def engine():
    """One engine (context) for the whole GPU session; fails loudly if the library or GPU is missing."""
    from theia_b200.engine import TadEngine
    eng = TadEngine(device=0)
    yield eng
    eng.close()
