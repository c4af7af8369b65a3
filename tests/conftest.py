"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"` tests: the oracle against the reference's known answers, host logic, C-ABI
symbol presence.  `-m gpu` tests: parity of the CUDA path (through the C-ABI) with the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def gl():
    """Gauss-Legendre nodes/weights generator (stands in for FastGaussQuadrature.gausslegendre;
    nodes/weights are inputs to both the oracle and the engine, so any accurate generator works)."""
    import numpy as np

    def _gl(n):
        return np.polynomial.legendre.leggauss(n)
    return _gl
