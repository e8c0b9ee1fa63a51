"""The CPU oracle against the reference's own saved MATLAB workspaces (tests/golden/*.npz, made by
tests/golden/make_golden.py from /root/reference/data) -- this is what pins the oracle."""
import numpy as np
import pytest


def test_model_matrices_bit_exact(orc, golden):
    g = golden["kat_matrices"]
    A, Av, A0, D = orc.model_mats(float(g["h"]), int(g["k_hor"]))
    assert np.array_equal(A, g["A"])            # getPosMat.m
    assert np.array_equal(A, g["A_p_dmpc"])     # dmpc_soft_bound.m:92-108
    assert np.array_equal(Av, g["A_v_dmpc"])
    assert np.array_equal(A0, g["A_initp"])
    assert np.array_equal(D, g["Delta"])        # getDeltaMat.m


def test_numpy_restatement_agrees(orc):
    for h, K in ((0.2, 15), (0.1, 20), (0.25, 7)):
        c = orc.model_mats(h, K)
        n = orc.np_model_mats(h, K)
        for a, b in zip(c, n):
            assert np.array_equal(a, b)


def _kat_step(orc, g, variant):
    P = orc.default_params(variant)
    assert P.K == int(g["k_hor"]) and P.h == float(g["h"]) and P.rmin == float(g["rmin"])
    assert P.c == float(g["c"]) and P.alim == float(g["alim"]) and P.term == float(g["term"])
    return P, orc.step(P, g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"], g["pmax"],
                       want_diag=True)


def test_single_step_kat_soft_bound(orc, golden):
    """data/failure_rate/failure_rate2.mat: N=200, step 14, agents 1..169 solved by solveSoftDMPCbound;
    agent 170 returned coll=1.  Tolerances of SURVEY section 8(d)(ii)."""
    g = golden["kat_soft_bound"]
    P, o = _kat_step(orc, g, orc.VARIANT_SOFT_BOUND)
    ns = int(g["n_solved"])
    assert o["first_fail"] == ns                      # the agent MATLAB stopped at
    assert o["status"][ns] & orc.ST_COLL
    assert np.all(o["status"][:ns] & orc.ST_SOLVED)
    err = np.abs(o["l_new"][:, :, :ns] - g["new_l"]).max(axis=(0, 1))
    kstar = np.array([o["diag"][n].kstar for n in range(ns)])
    free, coll = kstar == 0, kstar > 0
    assert free.sum() == 41 and coll.sum() == 128
    assert err[free].max() <= 1e-4                    # observed 5e-6
    assert err[coll].max() <= 1e-2                    # MATLAB ConstraintTolerance 1e-3 noise; observed 5.3e-3
    assert np.median(err[coll]) <= 1e-4
    # first columns = the applied state
    assert np.abs(o["p1"][:, :ns] - g["pk_new"]).max() <= 1e-3
    assert np.abs(o["a1"][:, :ns] - g["ak_new"]).max() <= 5e-2
    assert np.abs(o["a1"][:, :ns][:, free] - g["ak_new"][:, free]).max() <= 2e-3


def test_single_step_kat_soft_bound2(orc, golden):
    """data/comp_kctr/comp_kctr_3.mat: solveSoftDMPCbound2 (k_ctr = k-1), 9 solved agents."""
    g = golden["kat_soft_bound2"]
    P, o = _kat_step(orc, g, orc.VARIANT_SOFT_BOUND2)
    ns = int(g["n_solved"])
    assert np.all(o["status"][:ns] & orc.ST_SOLVED)
    err = np.abs(o["l_new"][:, :, :ns] - g["new_l"]).max(axis=(0, 1))
    assert err.max() <= 2e-4                          # observed 8.4e-5
    # the other variant must NOT fit (the fixture discriminates k_ctr)
    P1 = orc.default_params(orc.VARIANT_SOFT_BOUND)
    o1 = orc.step(P1, g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"], g["pmax"])
    err1 = np.abs(o1["l_new"][:, :, :ns] - g["new_l"]).max(axis=(0, 1))
    assert err1.max() > 1e-2


def test_kkt_certificates(orc, golden):
    g = golden["kat_soft_bound"]
    P, o = _kat_step(orc, g, orc.VARIANT_SOFT_BOUND)
    for n in range(int(g["N"])):
        if o["status"][n] & orc.ST_SOLVED:
            d = o["diag"][n]
            assert d.kkt_stat < 1e-9 and d.kkt_prim < 1e-9 and d.kkt_dual < 1e-7 and d.kkt_comp < 1e-6


def test_gi_against_independent_pdip(orc, golden):
    """the oracle's dual active-set solver vs an independent dense interior-point method on the
    dense QP assembled like solveSoftDMPCbound.m:60-98"""
    g = golden["kat_soft_bound"]
    P = orc.default_params(orc.VARIANT_SOFT_BOUND)
    checked = 0
    for n in (0, 3, 7, 20, 50, 99, 140):
        q = orc.np_dense_qp(P, g["pk_prev"][:, n], g["pf"][:, n], g["vk_prev"][:, n], g["ak_prev"][:, n], n,
                            g["l"], g["pmin"], g["pmax"])
        if q is None:
            continue
        H, f, A, b, lb, ub = q
        rc, x, lam, it, _ = orc.qp_gi(H, f, A, b, lb, ub)
        assert rc == 0
        scale = np.abs(f).max()
        x2, info = orc.qp_pdip(H / scale, f / scale, A, b, lb, ub)
        assert info["converged"]
        assert np.abs(x[:45] - x2[:45]).max() < 1e-6
        checked += 1
    assert checked >= 5


@pytest.mark.parametrize("variant", [2, 3])
def test_hard_variants_selfconsistent(orc, golden, variant):
    """solveHardDMPC / solveHardDMPCOnDemand: no saved reference data exists (parity unpinned);
    check the oracle's own optimality certificate and that solved agents respect the rows."""
    g = golden["kat_soft_bound"]
    P = orc.default_params(variant)
    o = orc.step(P, g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"], g["pmax"],
                 n0=0, n1=60, want_diag=True)
    solved = [n for n in range(60) if o["status"][n] & orc.ST_SOLVED]
    assert len(solved) > 20
    for n in solved:
        d = o["diag"][n]
        assert d.kkt_stat < 1e-9 and d.kkt_prim < 1e-9
        assert np.abs(o["l_new"][:, :, n]).max() < 50


def test_closed_loop_first_steps_vs_reference(orc, golden):
    """data/failure_rate/failure_rate3.mat: a complete N=200 MATLAB transition; raw accelerations
    recovered as ak/r_factor (stored as float32).  The first solved step reproduces
    MATLAB for ALL agents; afterwards discrete decisions on MATLAB's 1e-4 solver noise make single
    agents diverge (0.44 m/s^2 at step 2), so later steps are compared by quantile."""
    g = golden["ref_transition_n200"]
    P = orc.default_params(orc.VARIANT_SOFT_BOUND)
    r = orc.simulate(P, g["po"], g["pf"], g["pmin"], g["pmax"], max_steps=6, nthreads=4, stop_on_fail=False)
    a_ref = g["a_raw"].astype(float)
    assert np.abs(a_ref[:, 1, :] - r["ak"][:, 1, :]).max() <= 1e-5     # observed 3.7e-7 (float32 storage)
    for k in (2, 3, 4):
        e = np.abs(a_ref[:, k, :] - r["ak"][:, k, :]).max(0)
        assert np.quantile(e, 0.9) <= 1e-3, (k, np.quantile(e, 0.9))
    assert np.abs(r["ak"]).max() <= P.alim + 1e-9


def test_postprocess_matches_matlab_workspace(orc, golden):
    """failure_rate.m:134-195 on the finished N=200 trial of failure_rate3.mat: time scaling bit-exact,
    100 Hz not-a-knot splines to 1e-14, the trial's figures (collision flag, trajectory time) identical"""
    from tests.conftest import raw_transition
    g = golden["postprocess_n200"]
    pk, vk, ak = raw_transition(g)
    o = orc.postprocess(pk, vk, ak, g["pf"], float(g["h"]), c=float(g["c"]), rmin=float(g["rmin"]))
    assert o["r_factor"] == float(g["r_factor"]) and o["h_scaled"] == float(g["h_scaled"])
    assert o["nt"] == int(g["nt"]) and abs(o["T"] - float(g["T"])) < 1e-12
    sel = g["sel"]
    for k in ("pk", "vk", "ak"):
        assert np.array_equal(o[k][:, :, sel], g[k]), k
    for k in ("p", "v", "a"):
        assert np.abs(o[k][:, :, sel] - g[k]).max() < 1e-14, k
    assert np.array_equal(o["time_index"], g["time_index"])
    assert o["traj_time"] == float(g["traj_time"]) and o["violation"] == int(g["violation"]) == 1
    assert abs(o["totdist"] - float(g["totdist"])) < 1e-9
    # the spline is MATLAB's: cross-check the pure-numpy implementation against scipy's not-a-knot
    from scipy.interpolate import CubicSpline
    ref = CubicSpline(o["tk"], o["pk"][:, :, 3], axis=1, bc_type="not-a-knot")(o["t"])
    assert np.abs(ref - o["p"][:, :, 3]).max() < 1e-12
