import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built library (binaries stay out of history): build it once, like
    # __graft_entry__.build().  Without nvcc (a plain CPU box) the oracle / golden / gloo tests still run; the tests
    # that load the library fail on their own with the loader's message.
    from eas_snn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        try:
            from eas_snn_b200.build import build_library
            build_library()
        except (OSError, RuntimeError) as e:
            sys.stderr.write("conftest: could not build libeas_b200.so (%s)\n" % e)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
