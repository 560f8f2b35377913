import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "cl1226_golden.npz"))


@pytest.fixture(scope="session")
def cl1226_fit():
    """The shipped cluster rebuilt from the committed raw-input fixture (no /root/reference needed)."""
    from joxsz_b200 import cluster
    from joxsz_b200.mb import mb
    mb.fit.debugfit = False
    inp = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
    fit, sz = cluster.build_fit(inp, savedir=None)
    return fit


@pytest.fixture(scope="session")
def cl1226_oracle(cl1226_fit):
    from helpers import oracle_setup_from_fit
    return oracle_setup_from_fit(cl1226_fit)


@pytest.fixture(scope="session")
def cl1226_fit_integ():
    """Same cluster with the integrated-Compton-parameter penalty on (reference ``calc_integ = True``)."""
    from joxsz_b200 import cluster
    from joxsz_b200.mb import mb
    mb.fit.debugfit = False
    inp = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
    inp.calc_integ = True
    fit, sz = cluster.build_fit(inp, savedir=None)
    return fit
