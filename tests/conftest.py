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
    from oracle import binding
    return binding.load_oracle()


@pytest.fixture(scope="session")
def verbatim():
    from oracle import binding
    lib = binding.load_verbatim()
    if lib is None:
        pytest.skip("oracle/_ref/libref_verbatim.so not built (no /root/reference here)")
    return lib
