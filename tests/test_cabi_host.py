"""C-ABI library: loads, exports every symbol include/dmpc_b200.h declares, host-side entry points
work, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dmpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmpcb200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from multiagent_planning_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/dmpc_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names          # ctypes prototypes cover the whole header
    assert L.dmpcb200_abi_version() == 2


def test_default_params_are_the_reference_values():
    from multiagent_planning_b200 import dmpc
    P = dmpc.default_params(dmpc.SOFT_BOUND)
    # test/failure_rate.m:7-27, solveSoftDMPCbound.m:43-52,78
    assert (P.K, P.h, P.rmin, P.c, P.alim) == (15, 0.2, 0.35, 2.0, 1.0)
    assert (P.Q1, P.S1, P.term, P.Q_far, P.Q_near, P.S_free) == (1000.0, 100.0, -5e4, 1000.0, 10000.0, 10.0)
    assert (P.slack_lb, P.neigh_factor, P.coll_tol, P.max_tries) == (-0.05, 3.0, 0.05, 30)
    assert dmpc.default_params(dmpc.SOFT_BOUND2).slack_lb == -0.01   # solveSoftDMPCbound2.m:77
    with pytest.raises(TypeError):
        dmpc.default_params(0, nonsense=1)


def test_params_layout_matches_oracle_and_emul(orc):
    """the three ctypes mirrors of the parameter struct agree field by field"""
    from multiagent_planning_b200 import _lib
    from tests.host_emul import emul
    a = [(n, t) for n, t in _lib.Params._fields_]
    assert a == list(emul.Params._fields_)
    assert a[:-1] == list(orc.Params._fields_)     # the oracle has no goal_tol


def test_model_mats_host_entry_bit_exact(golden):
    """getPosMat.m / getDeltaMat.m / dmpc_soft_bound.m:81-108 through the C-ABI (host computation)"""
    from multiagent_planning_b200 import dmpc
    g = golden["kat_matrices"]
    A, Av, A0, D = dmpc.modelMats(float(g["h"]), int(g["k_hor"]))
    assert np.array_equal(A, g["A"]) and np.array_equal(Av, g["A_v_dmpc"])
    assert np.array_equal(A0, g["A_initp"]) and np.array_equal(D, g["Delta"])
    assert np.array_equal(dmpc.getPosMat(0.2, 15), g["A"])
    assert np.array_equal(dmpc.getDeltaMat(15), g["Delta"])
    with pytest.raises(dmpc.DmpcError):
        dmpc.getPosMat(0.2, 99)


def test_compute_fails_loudly_without_gpu():
    from multiagent_planning_b200 import dmpc
    if dmpc.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(dmpc.DmpcError, match="no CUDA device"):
        dmpc.Solver(8)
    with pytest.raises(dmpc.DmpcError):
        dmpc.initDMPC([0, 0, 1], [1, 1, 1], 0.2, 15)


def test_product_never_uses_the_oracle():
    """the oracle is test infrastructure: nothing in the package may import, link or include it"""
    pkg = os.path.join(ROOT, "multiagent_planning_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|liboracle|dmpc_oracle|host_emul|libemul")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not pat.search(txt), f"{f} references the oracle / test emulation"


def test_scenarios_respect_min_distance():
    from multiagent_planning_b200 import scenarios
    pmin, pmax = scenarios.density_arena(120)
    po, pf = scenarios.random_test(120, pmin, pmax, 0.35, 2.0, seed=3)
    for pts in (po, pf):
        assert (pts >= pmin[:, None]).all() and (pts <= pmax[:, None]).all()
        d = pts[:, :, None] - pts[:, None, :]
        d[2] /= 2.0
        dist = np.sqrt((d ** 2).sum(0)) + 10 * np.eye(120)
        assert dist.min() > 0.35
    po2, _ = scenarios.random_test(120, pmin, pmax, 0.35, 2.0, seed=3)
    assert np.array_equal(po, po2)


def test_mex_gateway_syntax_against_the_header():
    """matlab/dmpc_b200_mex.cpp cannot be built here (no MATLAB): syntax-check it against
    include/dmpc_b200.h with a stub mex.h so that the gateway cannot drift from the C-ABI."""
    import subprocess
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "mex_stub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "matlab", "dmpc_b200_mex.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_status_to_reference_flags():
    """status word -> [success/feasible, outbound, coll] per variant (solveSoftDMPCbound.m:17-19,29-30,125-128;
    solveSoftDMPCbound2.m:14,26,123-126; solveHardDMPC.m:14,76-79)"""
    from multiagent_planning_b200 import dmpc as d
    f = d.status_flags
    assert f(d.ST_SOLVED, 0) == (True, 1, 0, 0)
    assert f(d.ST_SOLVED | d.ST_OUTBOUND, 0) == (True, 1, 1, 0)      # bound: feasible stays 1
    assert f(d.ST_COLL, 0) == (False, 1, 0, 1)                       # bound: coll = 1, feasible = 1
    assert f(d.ST_INFEASIBLE | (30 << 8), 0) == (False, 0, 0, 0)
    assert f(d.ST_QPFAIL | d.ST_OVERFLOW, 0) == (False, 0, 0, 0)
    for v in (1, 2, 3):
        assert f(d.ST_SOLVED, v) == (True, 1, 0, 0)
        assert f(d.ST_SOLVED | d.ST_OUTBOUND, v) == (True, 0, 1, 0)  # success = 0; outbound = 1
        assert f(d.ST_COLL, v) == (False, 0, 0, 1)                   # success stays 0 on the collision exit
        assert f(d.ST_INFEASIBLE, v) == (False, 0, 0, 0)
