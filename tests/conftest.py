import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")
    # Built artefacts are git-ignored: make sure the C-ABI library (nvcc, sm_100a — cross-compiles without a GPU) and
    # the oracle exist before any test imports them.  `make` is a no-op when they are up to date.
    from spenso_b200 import _capi

    _capi.build_library()
    import _oracle

    _oracle.build_oracle()
