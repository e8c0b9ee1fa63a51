import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def orc():
    from oracle import dmpc_oracle
    dmpc_oracle.build()
    return dmpc_oracle


@pytest.fixture(scope="session")
def emul():
    from tests.host_emul import emul as e
    e.build()
    return e


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {n: np.load(os.path.join(d, n + ".npz")) for n in
            ("kat_matrices", "kat_soft_bound", "kat_soft_bound2", "ref_transition_n200")}


@pytest.fixture(scope="session")
def dmpc():
    """The product package; GPU tests fail loudly (not skip) when the library is missing."""
    from multiagent_planning_b200 import _lib, dmpc as d
    _lib.lib()
    return d


def oracle_params(orc, P):
    """oracle Params with the same values as a product Params"""
    O = orc.default_params(P.variant)
    for n, _ in O._fields_:
        setattr(O, n, getattr(P, n))
    return O
