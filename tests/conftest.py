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
            ("kat_matrices", "kat_soft_bound", "kat_soft_bound2", "ref_transition_n200", "postprocess_n200")}


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


def raw_transition(g):
    """raw MPC trajectory (pk, vk, ak of the loop, (3, S, N)) of the golden post-processing fixture: v and p
    follow from the stored raw accelerations and start points by the model recursion"""
    import numpy as np
    a, po, h = g["ak_raw"], g["po"], float(g["h"])
    S = a.shape[1]
    v, p = np.zeros_like(a), np.zeros_like(a)
    p[:, 0, :] = po
    for k in range(1, S):
        v[:, k, :] = v[:, k - 1, :] + h * a[:, k, :]
        p[:, k, :] = p[:, k - 1, :] + h * v[:, k - 1, :] + h * h / 2 * a[:, k, :]
    return p, v, a
