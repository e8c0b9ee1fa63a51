"""GPU parity tests: the CUDA path (through the C-ABI of libdmpc_b200.so) against the CPU oracle on
the same inputs, against the reference's golden MATLAB workspaces, and -- at the full sizes of
BASELINE.json -- through size-independent properties.  Tolerances:
  GPU vs oracle, teacher-forced: max |dp| over the horizon <= 1e-6 m, identical status flags;
  GPU vs MATLAB golden: collision-free agents <= 1e-4 m, colliding agents <= 1e-2 m."""
import os

import numpy as np
import pytest

from tests.conftest import oracle_params

pytestmark = pytest.mark.gpu

TOL = 1e-6


def _solver(dmpc, g, variant, **kw):
    P = dmpc.default_params(variant, **kw)
    N = g["l"].shape[2]
    return P, dmpc.Solver(N, P, pmin=g["pmin"], pmax=g["pmax"], pf=g["pf"])


def _cmp_step(orc, P, s, pk, vk, ak, pf, l, pmin, pmax, tol=TOL):
    g = s.step(pk, vk, ak, l, want_horizons=True)
    o = orc.step(oracle_params(orc, P), pk, vk, ak, pf, l, pmin, pmax)
    n0, n1 = s.n0, s.n1
    # flags (bits 0..7) AND the number of infeasible-retries taken (bits 8..15): the Farkas short-cut of the
    # retry loop must land on the reference's try count, not just on the same solution
    assert np.array_equal(g["status"][n0:n1] & 0xFFFF, o["status"][n0:n1] & 0xFFFF)
    for k in ("l_new", "p1", "v1", "a1"):
        assert np.abs(g[k] - o[k]).max() <= tol, k
    assert g["first_fail"] == o["first_fail"]
    return g, o


def test_kat_soft_bound_vs_oracle_and_matlab(dmpc, orc, golden):
    g = golden["kat_soft_bound"]
    P, s = _solver(dmpc, g, dmpc.SOFT_BOUND)
    with s:
        out, o = _cmp_step(orc, P, s, g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"], g["pmax"])
        ns = int(g["n_solved"])
        assert out["first_fail"] == ns and out["status"][ns] & dmpc.ST_COLL
        err = np.abs(out["l_new"][:, :, :ns] - g["new_l"]).max(axis=(0, 1))
        free = out["diag"]["kstar"][:ns] == 0
        assert free.sum() == 41
        assert err[free].max() <= 1e-4 and err[~free].max() <= 1e-2 and np.median(err[~free]) <= 1e-4
        assert np.abs(out["p1"][:, :ns] - g["pk_new"]).max() <= 1e-3
        # v, a horizons are consistent with the model: p = A_p a + A_initp [po;vo]
        A, Av, A0, _ = dmpc.modelMats(P.h, P.K)
        for n in (0, 5, 100):
            a = out["a_hor"][:, :, n].reshape(-1, order="F")
            p = A @ a + A0 @ np.r_[g["pk_prev"][:, n], g["vk_prev"][:, n]]
            v = Av @ a + np.tile(g["vk_prev"][:, n], P.K)
            assert np.abs(p - out["l_new"][:, :, n].reshape(-1, order="F")).max() < 1e-12
            assert np.abs(v - out["v_hor"][:, :, n].reshape(-1, order="F")).max() < 1e-12


def test_kat_soft_bound2_vs_oracle_and_matlab(dmpc, orc, golden):
    g = golden["kat_soft_bound2"]
    P, s = _solver(dmpc, g, dmpc.SOFT_BOUND2)
    with s:
        out, _ = _cmp_step(orc, P, s, g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"], g["pmax"])
        ns = int(g["n_solved"])
        assert np.abs(out["l_new"][:, :, :ns] - g["new_l"]).max() <= 2e-4


@pytest.mark.parametrize("variant", [2, 3])
def test_hard_variants_vs_oracle(dmpc, orc, golden, variant):
    g = golden["kat_soft_bound2"]    # N = 100 (BASELINE config 2 shape)
    P, s = _solver(dmpc, g, variant)
    with s:
        _cmp_step(orc, P, s, g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"], g["pmax"])


def test_cpp_neighbour_mode_vs_oracle(dmpc, orc, golden):
    """C++ check_collisionsv2 threshold rmin*(1+k/K) (dmpc.cpp:418) as a flag"""
    g = golden["kat_soft_bound"]
    P, s = _solver(dmpc, g, dmpc.SOFT_BOUND, neigh_mode=1)
    with s:
        _cmp_step(orc, P, s, g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"], g["pmax"])


def test_matlab_named_surface(dmpc, orc, golden):
    """solveSoftDMPCbound / CheckCollSoftDMPC / CollConstrSoftDMPC / propStatedmpc / initDMPC /
    ReachedGoal with the reference's argument lists (1-based n, k)."""
    g = golden["kat_soft_bound"]
    l, K, h = g["l"], 15, 0.2
    A, Av, A0, Delta = dmpc.modelMats(h, K)
    E1, E2 = np.diag([1, 1, 1 / 2.0]), np.diag([1, 1, 1 / 4.0])
    O = orc.default_params(0)
    for n in (1, 2, 17, 120):
        i = n - 1
        p, v, a, feas, outb, coll = dmpc.solveSoftDMPCbound(
            g["pk_prev"][:, i], g["pf"][:, i], g["vk_prev"][:, i], g["ak_prev"][:, i], n, h, l, K, 0.35,
            g["pmin"], g["pmax"], 1.0, A, A0, A, Av, Delta, 1000, 100, E1, E2, 2, -5e4)
        assert (feas, outb, coll) == (1, 0, 0) and p.shape == (3, K)
        assert np.abs(p - g["new_l"][:, :, i]).max() <= 1e-2
        st, po_, vo_, ao_, _ = orc.solve_agent(O, g["pk_prev"][:, i], g["pf"][:, i], g["vk_prev"][:, i],
                                               g["ak_prev"][:, i], i, l, g["pmin"], g["pmax"])
        assert np.abs(p - po_).max() <= TOL and np.abs(v - vo_).max() <= TOL and np.abs(a - ao_).max() <= TOL
        p2, v2 = dmpc.propStatedmpc(g["pk_prev"][:, i], g["vk_prev"][:, i], a.reshape(-1, order="F"), A0, A, Av)
        assert np.abs(p2 - p.reshape(-1, order="F")).max() < 1e-12
        assert np.abs(v2 - v.reshape(-1, order="F")).max() < 1e-12
    # agent 170: k == 1 violation -> coll = 1 and empty outputs (solveSoftDMPCbound.m:25-32)
    n = int(g["n_failed"])
    p, v, a, feas, outb, coll = dmpc.solveSoftDMPCbound(
        g["pk_prev"][:, n - 1], g["pf"][:, n - 1], g["vk_prev"][:, n - 1], g["ak_prev"][:, n - 1], n, h, l, K, 0.35,
        g["pmin"], g["pmax"], 1.0, A, A0, A, Av, Delta, 1000, 100, E1, E2, 2, -5e4)
    assert coll == 1 and feas == 1 and p.size == 0
    # CheckColl / CollConstr on agent 3 at every horizon step
    n = 3
    for k in range(1, K + 1):
        p3 = l[:, k - 1, n - 1]
        viol, md, vc = dmpc.CheckCollSoftDMPC(p3, l, n, k, E1, 0.35, 2)
        oany, ov, ovc, omd = orc.check_coll(O, p3, l, n - 1, k)
        assert np.array_equal(viol, ov.astype(bool)) and np.array_equal(vc, ovc.astype(bool)) and md == omd
        if vc.any():
            Ain, b, pd = dmpc.CollConstrSoftDMPC(p3, g["pk_prev"][:, n - 1], g["vk_prev"][:, n - 1], n, k, l, 0.35,
                                                 None, A0, E1, E2, 2, vc)
            oA, ob, opd, _ = orc.coll_constr(O, p3, g["pk_prev"][:, n - 1], g["vk_prev"][:, n - 1], n - 1, k, l,
                                             mask=vc.astype(np.uint8))
            assert Ain.shape == oA.shape == (int(vc.sum()), 3 * K)
            assert np.abs(Ain - oA).max() < 1e-13 and np.abs(b - ob).max() < 1e-12 and np.array_equal(pd, opd)
    with pytest.raises(dmpc.DmpcError):
        dmpc.CheckCollSoftDMPC(l[:, 0, 0], l, 1, 1, E1, 0.35, 4)      # order 4 is rejected
    # initDMPC / ReachedGoal
    p, v, a = dmpc.initDMPC(g["pk_prev"][:, 0], g["pf"][:, 0], h, K, 150)
    op, _, _ = orc.init_dmpc(g["pk_prev"][:, 0], g["pf"][:, 0], h, K, 10.0)
    assert np.array_equal(p, op) and not v.any() and not a.any()
    pk = np.repeat(g["pf"][:, None, :], 2, axis=1) + 0.003
    assert dmpc.ReachedGoal(pk, g["pf"], 2, 0.01, 200) is True
    pk[0, 1, 7] += 0.02
    assert dmpc.ReachedGoal(pk, g["pf"], 2, 0.01, 200) is False
    dmpc.close_cached_solvers()


def test_hard_constraint_helpers_and_dec_iscp_names(dmpc, orc, golden):
    """CollConstrHardDMPC / CollConstrHardDMPCOnDemand at helper level (reference shapes: N-1 rows with
    all-zero padding, CollConstrHardDMPC.m:3-5), and the dec-iSCP-named helpers CollConstr / propState
    against a numpy restatement of dec-iSCP/CollConstr.m:1-23 and propState.m:1-10."""
    g = golden["kat_soft_bound2"]
    l, K, h, N = g["l"], 15, 0.2, 100
    A, Av, A0, _ = dmpc.modelMats(h, K)
    E1, E2 = np.diag([1, 1, 1 / 2.0]), np.diag([1, 1, 1 / 4.0])
    OH, OD = orc.default_params(orc.VARIANT_HARD), orc.default_params(3)
    n = 5
    po, vo = g["pk_prev"][:, n - 1], g["vk_prev"][:, n - 1]
    seen = 0
    for k in (1, 4, 9, 15):
        p3 = l[:, k - 1, n - 1]
        Ain, b = dmpc.CollConstrHardDMPC(p3, po, vo, n, k, l, 0.35, A, A0, E1, E2, 2)
        oA, ob, _, _ = orc.coll_constr(OH, p3, po, vo, n - 1, k, l)
        nr = oA.shape[0]
        seen += nr
        assert Ain.shape == (N - 1, 3 * K) and b.shape == (N - 1,)
        assert np.abs(Ain[:nr] - oA).max() < 1e-13 and np.abs(b[:nr] - ob).max() < 1e-12
        assert not Ain[nr:].any() and not b[nr:].any()                      # the reference's vacuous rows
        _, _, vc = dmpc.CheckCollSoftDMPC(p3, l, n, k, E1, 0.35, 2)
        if vc.any():
            Ad, bd = dmpc.CollConstrHardDMPCOnDemand(p3, po, vo, n, k, l, 0.35, A, A0, E1, E2, 2, vc)
            oA, ob, _, _ = orc.coll_constr(OD, p3, po, vo, n - 1, k, l, mask=vc.astype(np.uint8))
            assert Ad.shape == oA.shape and np.abs(Ad - oA).max() < 1e-13 and np.abs(bd - ob).max() < 1e-12
    assert seen > 0
    # dec-iSCP/CollConstr.m: all columns of l are obstacles, row on block k-1, no velocity term
    obs = l[:, :, 10:16]
    p3, k = l[:, 6, 3] + 0.01, 7
    Ac, bc = dmpc.CollConstr(p3, po, k, obs, A, 0.35, E1, E2, 2)
    assert Ac.shape == (6, 3 * K)
    for i in range(6):
        d = p3 - obs[:, k - 1, i]
        dist = np.linalg.norm(E1 @ d)
        diff = E2 @ d
        r = dist * (0.35 - dist + diff @ p3 / dist) - diff @ po
        row = np.zeros(3 * K)
        row[3 * (k - 2):3 * (k - 1)] = diff
        assert np.abs(Ac[i] - (-row @ A)).max() < 1e-12 and abs(bc[i] + r) < 1e-12
    # dec-iSCP/propState.m
    a = np.random.default_rng(3).uniform(-1, 1, 3 * K)
    pp, vv = dmpc.propState(po, a, A[:3 * (K - 1)], Av[:3 * (K - 1)], K)
    assert np.abs(pp - np.r_[po, A[:3 * (K - 1)] @ a + np.tile(po, K - 1)]).max() < 1e-12
    assert np.abs(vv - np.r_[np.zeros(3), Av[:3 * (K - 1)] @ a]).max() < 1e-12
    dmpc.close_cached_solvers()


def test_drop_ins_do_not_touch_the_resident_state(dmpc, golden):
    """solve_agent / check_coll / coll_constr run on the handle's scratch: a per-agent call between
    init_horizons and run must not change the closed loop (the MEX gateway shares one handle)"""
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config("N100")
    P = dmpc.default_params(0)
    g = golden["kat_soft_bound2"]
    with dmpc.Solver(100, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        s.init_horizons(cfg["po"])
        ref = s.run(12, record=True)
        s.init_horizons(cfg["po"])
        s.run(5)
        s.solve_agent(g["pk_prev"][:, 7], g["pf"][:, 7], g["vk_prev"][:, 7], g["ak_prev"][:, 7], 7, g["l"])
        s.check_coll(g["l"][:, 2, 7], g["l"], 7, 3)
        s.coll_constr(g["l"][:, 2, 7], g["pk_prev"][:, 7], g["vk_prev"][:, 7], 7, 3, g["l"], mask=np.ones(100, np.uint8))
        r = s.run(7, record=True)
        assert np.array_equal(r["pk"][:, -1, :], ref["pk"][:, 12, :])
        assert np.array_equal(s.get_state()["pk"], ref["pk"][:, 12, :])


def test_cpp_semantics_preset_and_static_obstacles_vs_oracle(dmpc, orc):
    """The C++ port's semantics as flags (dmpcb200_default_params_cpp: neighbour threshold rmin (1 + k/K),
    slack bound 0.01 doubled for <= 20 retries, term -1e6, k_ctr = k + k_factor) and its un-commanded agents
    (N_cmd < N: static obstacles, dmpc.cpp:1633-1649), teacher-forced against the oracle run with the same
    parameters on agents [0, N_cmd)."""
    import ctypes as C
    from multiagent_planning_b200 import _lib, scenarios
    N, n_cmd = 60, 45
    pmin, pmax = scenarios.density_arena(N, density=2.5)
    po, pf = scenarios.random_test(N, pmin, pmax, 0.5, 1.5, seed=31)
    pf[:, n_cmd:] = po[:, n_cmd:]
    for k_factor in (0, -1):
        P = _lib.Params()
        _lib.lib().dmpcb200_default_params_cpp(C.byref(P), k_factor)
        P.K = 15
        assert P.variant == (1 if k_factor else 0) and P.neigh_mode == 1 and P.max_tries == 21
        O = oracle_params(orc, P)
        with dmpc.Solver(N, P, pmin=pmin, pmax=pmax, pf=pf) as s:
            s.set_static_obstacles(n_cmd)
            l, pk, vk, ak = s.init_horizons(po)
            assert np.array_equal(l[:, :, n_cmd:], np.repeat(po[:, None, n_cmd:], 15, axis=1))   # constant horizons
            retried = 0
            for _ in range(14):
                g = s.step(pk, vk, ak, l)
                o = orc.step(O, pk, vk, ak, pf, l, pmin, pmax, n1=n_cmd)
                assert np.array_equal(g["status"][:n_cmd] & 0xFFFF, o["status"][:n_cmd] & 0xFFFF)
                assert np.abs(g["l_new"] - o["l_new"]).max() <= TOL and np.abs(g["a1"] - o["a1"]).max() <= TOL
                assert np.array_equal(g["l_new"][:, :, n_cmd:], l[:, :, n_cmd:])                   # obstacles untouched
                retried += int((((o["status"][:n_cmd] >> 8) & 0xFF) > 0).sum())
                l, pk, vk, ak = g["l_new"], g["p1"], g["v1"], g["a1"]
            # the resident loop: obstacles stay, the goal test covers the commanded agents only
            s.init_horizons(po)
            r = s.run(120, record=True)
            st = s.get_state()
            assert np.array_equal(st["pk"][:, n_cmd:], po[:, n_cmd:]) and np.array_equal(st["l"][:, 0, n_cmd:], po[:, n_cmd:])
            if r["reached"]:
                assert np.sqrt(((st["pk"][:, :n_cmd] - pf[:, :n_cmd]) ** 2).sum(0)).max() < P.goal_tol


def test_cpp_facade_solve_parallel(dmpc):
    """the Python mirror of class DMPC (dmpc/cpp/dmpc.h:70-182): solveParallelDMPCv2 == Solver.run"""
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config("N100")
    d = dmpc.DMPC(dict(T=20.0))
    d.set_boundaries(cfg["pmin"], cfg["pmax"])
    d.set_initial_pts(cfg["po"])
    d.set_final_pts(cfg["pf"])
    d.set_k_factor(0)
    d.set_cluster_num(8)
    sol = d.solveParallelDMPCv2(stop_on_fail=False)
    with dmpc.Solver(100, dmpc.default_params(0), pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        s.init_horizons(cfg["po"])
        r = s.run(99, record=True)
    assert len(sol) == 100 and d.steps == r["steps"] and d.successful == (r["reached"] and r["first_fail_step"] < 0)
    assert np.array_equal(sol[17]["pos"], r["pk"][:, :, 17]) and np.array_equal(sol[99]["acc"], r["ak"][:, :, 99])
    with pytest.raises(dmpc.DmpcError):
        d.set_k_factor(2)


def _closed_loop(dmpc, orc, cfg, steps, check_every=1):
    """device-resident run == host-stepped run; host-stepped teacher-forced vs oracle"""
    P = dmpc.default_params(cfg["variant"], **cfg["params"])
    N = cfg["N"]
    with dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        s.init_horizons(cfg["po"])
        r = s.run(steps, record=True, status_hist=True)
        l, pk, vk, ak = s.init_horizons(cfg["po"])
        assert np.array_equal(r["pk"][:, 0, :], pk)
        flips = 0
        for k in range(r["steps"]):
            if k % check_every == 0:
                g, o = s.step(pk, vk, ak, l), orc.step(oracle_params(orc, P), pk, vk, ak, cfg["pf"], l, cfg["pmin"],
                                                       cfg["pmax"], nthreads=4)
                assert np.array_equal(g["status"] & 0xFFFF, o["status"] & 0xFFFF)
                assert np.abs(g["l_new"] - o["l_new"]).max() <= TOL
            else:
                g = s.step(pk, vk, ak, l)
            assert np.array_equal(r["pk"][:, k + 1, :], g["p1"])       # same kernels, same bits
            assert np.array_equal(r["vk"][:, k + 1, :], g["v1"]) and np.array_equal(r["ak"][:, k + 1, :], g["a1"])
            assert np.array_equal(r["status_hist"][k], g["status"])
            l, pk, vk, ak = g["l_new"], g["p1"], g["v1"], g["a1"]
        # plain-launch and timed modes give the same trajectory as the graph
        for mode in (1, 2):
            s.init_horizons(cfg["po"])
            r2 = s.run(steps, mode=mode, record=True)
            assert r2["steps"] == r["steps"] and np.array_equal(r2["pk"], r["pk"])
        st = s.get_state()
        assert np.array_equal(st["pk"], pk) and np.array_equal(st["l"], l)
    return r


def test_closed_loop_4_agents_corner_swap(dmpc, orc):
    """BASELINE config 1 shape: dmpc_soft_bound.m:43-54, 100 fixed steps"""
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config("C1")
    r = _closed_loop(dmpc, orc, cfg, 100)
    end = r["pk"][:, -1, :]
    assert np.sqrt(((end - cfg["pf"]) ** 2).sum(0)).max() < 0.05
    assert r["first_fail_step"] == -1


def test_closed_loop_n100(dmpc, orc):
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config("N100")
    r = _closed_loop(dmpc, orc, cfg, 149, check_every=3)
    assert r["reached"] and r["steps"] < 120
    # early exit at the goal: the loop stopped by itself (ReachedGoal on the device)
    assert np.sqrt(((r["pk"][:, -1, :] - cfg["pf"]) ** 2).sum(0)).max() < 0.01


def test_closed_loop_hard_n100(dmpc, orc):
    """BASELINE config 2: N=100, 10x10x3 arena, solveHardDMPC"""
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config("C2")
    _closed_loop(dmpc, orc, cfg, 30, check_every=2)


def test_edge_cases(dmpc, orc):
    pmin, pmax = np.array([-2.5, -2.5, 0.2]), np.array([2.5, 2.5, 2.2])
    P = dmpc.default_params(0)
    O = oracle_params(orc, P)
    # one agent, two agents head-on, overlapping agents, a ragged tile (N = 33), N = 32, N = 65
    cases = {
        1: (np.array([[0.0], [0.0], [1.0]]), np.array([[1.0], [1.0], [1.5]])),
        2: (np.array([[-1.0, 1.0], [0.0, 0.01], [1.0, 1.0]]), np.array([[1.0, -1.0], [0.0, 0.0], [1.0, 3.5]])),
    }
    from multiagent_planning_b200 import scenarios
    for N in (3, 32, 33, 65):
        a, b = scenarios.density_arena(N, density=3.0)
        cases[N] = scenarios.random_test(N, a, b, 0.35, 2.0, seed=N) + (a, b)
    for N, c in cases.items():
        po, pf = c[0], c[1]
        lo, hi = (c[2], c[3]) if len(c) == 4 else (pmin, pmax)
        with dmpc.Solver(N, P, pmin=lo, pmax=hi, pf=pf) as s:
            l, pk, vk, ak = s.init_horizons(po)
            for n in range(N):
                assert np.array_equal(l[:, :, n], orc.init_dmpc(po[:, n], pf[:, n], P.h, P.K, 10.0)[0])
            for _ in range(12):
                g, _ = _cmp_step(orc, P, s, pk, vk, ak, pf, l, lo, hi)
                l, pk, vk, ak = g["l_new"], g["p1"], g["v1"], g["a1"]
    # overlap at k = 1 -> coll for both
    po = np.array([[0.0, 0.1], [0.0, 0.0], [1.0, 1.0]])
    pf = np.array([[1.0, -1.0], [0.0, 0.0], [1.0, 1.0]])
    with dmpc.Solver(2, P, pmin=pmin, pmax=pmax, pf=pf) as s:
        l, pk, vk, ak = s.init_horizons(po)
        g, _ = _cmp_step(orc, P, s, pk, vk, ak, pf, l, pmin, pmax)
        assert (g["status"] & dmpc.ST_COLL).all() and g["first_fail"] == 0
        assert np.array_equal(g["l_new"], l) and np.array_equal(g["p1"], pk)
    # API errors, not crashes
    with pytest.raises(dmpc.DmpcError):
        dmpc.Solver(0)
    with pytest.raises(dmpc.DmpcError):
        dmpc.Solver(4, dmpc.default_params(0, K=40))
    with dmpc.Solver(4, P) as s:
        with pytest.raises(dmpc.DmpcError):
            s.run(5)                      # no bounds / init yet


def test_sharded_handles_equal_whole(dmpc, orc, golden):
    """agents [n0,n1) per handle (the multi-GPU partition) give the rows of the whole solve"""
    g = golden["kat_soft_bound"]
    P, whole = _solver(dmpc, g, 0)
    with whole:
        w = whole.step(g["pk_prev"], g["vk_prev"], g["ak_prev"], g["l"])
    for n0, n1 in ((0, 63), (63, 126), (126, 200)):
        with dmpc.Solver(200, P, n0=n0, n1=n1, pmin=g["pmin"], pmax=g["pmax"], pf=g["pf"]) as s:
            o = s.step(g["pk_prev"], g["vk_prev"], g["ak_prev"], g["l"])
            assert np.array_equal(o["l_new"][:, :, n0:n1], w["l_new"][:, :, n0:n1])
            assert np.array_equal(o["status"][n0:n1], w["status"][n0:n1])
            assert np.array_equal(o["l_new"][:, :, :n0], g["l"][:, :, :n0])        # others untouched


def test_rescue_path_small_onchip_capacity(dmpc, orc, golden, monkeypatch):
    """force a tiny on-chip active-set capacity: agents that outgrow it re-solve in the global
    rescue slots and must still match the oracle"""
    g = golden["kat_soft_bound"]
    monkeypatch.setenv("DMPCB200_QMAX", "4")
    P, s = _solver(dmpc, g, 0)
    with s:
        assert s.config()["QMAX"] == 4
        _cmp_step(orc, P, s, g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"], g["pmax"])


def test_step_dev_with_caller_owned_unpadded_buffers(dmpc, orc, golden):
    """dmpcb200_step_dev on torch-owned device memory without tile padding (ragged last tile path)"""
    import torch
    g = golden["kat_soft_bound2"]
    N, K = 100, 15
    P, s = _solver(dmpc, g, 1)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)            # (3,N) -> (N,3)
    l_prev = torch.from_numpy(np.ascontiguousarray(g["l"].transpose(2, 1, 0))).to(dev)
    l_new = l_prev.clone()
    pk, vk, ak = t(g["pk_prev"]), t(g["vk_prev"]), t(g["ak_prev"])
    p1, v1, a1 = pk.clone(), vk.clone(), ak.clone()
    st = torch.zeros(N, dtype=torch.int32, device=dev)
    with s:
        s.step_dev(pk.data_ptr(), vk.data_ptr(), ak.data_ptr(), l_prev.data_ptr(), l_new.data_ptr(), p1.data_ptr(),
                   v1.data_ptr(), a1.data_ptr(), st.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        o = orc.step(oracle_params(orc, P), g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"],
                     g["pmax"])
        assert np.array_equal(st.cpu().numpy() & 0xFF, o["status"] & 0xFF)
        assert np.abs(l_new.cpu().numpy().transpose(2, 1, 0) - o["l_new"]).max() <= TOL
        assert np.abs(p1.cpu().numpy().T - o["p1"]).max() <= TOL
        out = torch.zeros(2, dtype=torch.float64, device=dev)
        s.goal_dev(l_new.data_ptr(), 3 * K, out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        md = np.sqrt(((l_new.cpu().numpy()[:, 0, :].T - g["pf"]) ** 2).sum(0)).max()
        assert abs(out.cpu().numpy()[0] - md) < 1e-12


# ---- full-size configurations: size-independent properties --------------------------------------
def _properties(dmpc, cfg, steps):
    P = dmpc.default_params(cfg["variant"], **cfg["params"])
    N, K = cfg["N"], P.K
    with dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        l, pk, vk, ak = s.init_horizons(cfg["po"])
        for _ in range(steps):
            g = s.step(pk, vk, ak, l, want_horizons=True)
            l, pk, vk, ak = g["l_new"], g["p1"], g["v1"], g["a1"]
        st = g["status"]
        ok = (st & dmpc.ST_SOLVED) != 0
        assert ok.mean() > 0.9 and not (st & (dmpc.ST_QPFAIL | dmpc.ST_OVERFLOW)).any()
        # feasibility of what was solved: acceleration box, workspace box
        assert np.abs(g["a_hor"][:, :, ok]).max() <= P.alim + 1e-9
        lo, hi = cfg["pmin"][:, None, None], cfg["pmax"][:, None, None]
        assert (g["l_new"][:, :, ok] >= lo - 1e-8).all() and (g["l_new"][:, :, ok] <= hi + 1e-8).all()
        # permutation equivariance: relabelling the agents permutes the result (unique optimum)
        rng = np.random.default_rng(0)
        perm = rng.permutation(N)
        g1 = s.step(pk, vk, ak, l)
    with dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"][:, perm]) as s2:
        g2 = s2.step(pk[:, perm], vk[:, perm], ak[:, perm], l[:, :, perm])
    assert np.array_equal(g1["status"][perm] & 0xFF, g2["status"] & 0xFF)
    assert np.abs(g1["l_new"][:, :, perm] - g2["l_new"]).max() <= 1e-8
    return g


def test_properties_n500_k15(dmpc):
    from multiagent_planning_b200 import scenarios
    _properties(dmpc, scenarios.config("C3"), 6)


def test_properties_n2000_k20_dense(dmpc):
    from multiagent_planning_b200 import scenarios
    _properties(dmpc, scenarios.config("C4"), 4)


def test_n500_vs_oracle_dense_steps(dmpc, orc):
    """the bench workload (N=500, K=15, soft bound): first steps teacher-forced against the oracle"""
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config("C3")
    P = dmpc.default_params(cfg["variant"])
    with dmpc.Solver(500, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        l, pk, vk, ak = s.init_horizons(cfg["po"])
        for _ in range(8):
            g, _ = _cmp_step(orc, P, s, pk, vk, ak, cfg["pf"], l, cfg["pmin"], cfg["pmax"])
            l, pk, vk, ak = g["l_new"], g["p1"], g["v1"], g["a1"]


def test_c4_n2000_k20_dense_vs_oracle(dmpc, orc):
    """BASELINE config 4 (N=2000, K=20, 2 agents/m^3): qp_kernel<4,20> (persistent grid) and scan_kernel<8,1,20>
    teacher-forced against the oracle, flags and retry counts included"""
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config("C4")
    P = dmpc.default_params(cfg["variant"], **cfg["params"])
    with dmpc.Solver(cfg["N"], P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        l, pk, vk, ak = s.init_horizons(cfg["po"])
        retried = 0
        for _ in range(5):
            g, o = _cmp_step(orc, P, s, pk, vk, ak, cfg["pf"], l, cfg["pmin"], cfg["pmax"])
            retried += int((((o["status"] >> 8) & 0xFF) > 0).sum())
            l, pk, vk, ak = g["l_new"], g["p1"], g["v1"], g["a1"]
        assert (g["diag"]["kstar"] > 0).sum() > 500      # it is the dense case


def test_n2000_k15_vs_oracle(dmpc, orc):
    """N = 2000, K = 15 (north_star's largest swarm): 14 dense steps teacher-forced against the oracle.  Step 3 holds
    an agent whose active set passes next to linear dependence (delta ~ 1e-7): without the multiplier check after the
    polish the solver stopped 2.9e-5 m from the optimum with a constraint active at a negative multiplier.  Step 13
    holds an agent whose first try is infeasible and whose active set becomes numerically dependent on the
    register-resident solver: it must go to the generic solver (rescue path) and come back with the oracle's retry
    count instead of reporting the inconsistent point as solved."""
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config("N2000")
    P = dmpc.default_params(cfg["variant"], **cfg["params"])
    with dmpc.Solver(cfg["N"], P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        l, pk, vk, ak = s.init_horizons(cfg["po"])
        for _ in range(14):
            g, o = _cmp_step(orc, P, s, pk, vk, ak, cfg["pf"], l, cfg["pmin"], cfg["pmax"])
            l, pk, vk, ak = o["l_new"], o["p1"], o["v1"], o["a1"]


@pytest.mark.parametrize("name,seed,steps", [("C3", 11, 16), ("C3", 21, 16), ("N2000", 24, 17)])
def test_bound2_thin_feasible_sets_vs_oracle(dmpc, orc, name, seed, steps):
    """solveSoftDMPCbound2 (rows one horizon index earlier: many retries, feasible sets that are merely thin).  The
    three cases the wide soak found: a try declared infeasible on the GPU although its feasible set is 5e-4 wide
    (C3 / seed 11, step 1: verdict after an ill-conditioned add), constraints called dependent at delta = 4e-10
    (seed 21, step 14; N = 2000 / seed 24, steps 3 and 16).  Flags AND retry counts must be the oracle's."""
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config(name)
    po, pf = scenarios.random_test(cfg["N"], cfg["pmin"], cfg["pmax"], 0.35, 2.0, seed)
    P = dmpc.default_params(dmpc.SOFT_BOUND2)
    with dmpc.Solver(cfg["N"], P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=pf) as s:
        l, pk, vk, ak = s.init_horizons(po)
        retried = 0
        for _ in range(steps):
            g, o = _cmp_step(orc, P, s, pk, vk, ak, pf, l, cfg["pmin"], cfg["pmax"])
            retried += int((((o["status"] >> 8) & 0xFF) > 0).sum())
            l, pk, vk, ak = o["l_new"], o["p1"], o["v1"], o["a1"]
        assert retried > 50


@pytest.mark.parametrize("N,density,seed,steps", [(300, 2.0, 12001, 10), (200, 3.0, 13058, 4)])
def test_bound2_ill_conditioned_polish_vs_oracle(dmpc, orc, N, density, seed, steps):
    """solveSoftDMPCbound2 on dense small swarms: the polish of an ill-conditioned (not inconsistent) active set
    stagnates at a residual of ~1e-9; judged at 1e-9 it was a reported solver failure where the oracle solves the try
    (host-build soak, seeds 12001 / step 8 / agent 73 and 13058 / step 2 / agent 175).  The line between noise and an
    inconsistent set (which leaves ~1e-2) is 1e-6.  Flags AND retry counts must be the oracle's."""
    from multiagent_planning_b200 import scenarios
    pmin, pmax = scenarios.density_arena(N, density)
    po, pf = scenarios.random_test(N, pmin, pmax, 0.35, 2.0, seed)
    P = dmpc.default_params(dmpc.SOFT_BOUND2)
    with dmpc.Solver(N, P, pmin=pmin, pmax=pmax, pf=pf) as s:
        l, pk, vk, ak = s.init_horizons(po)
        for _ in range(steps):
            g, o = _cmp_step(orc, P, s, pk, vk, ak, pf, l, pmin, pmax)
            l, pk, vk, ak = o["l_new"], o["p1"], o["v1"], o["a1"]


def test_k20_small_swarm_vs_oracle(dmpc, orc):
    """K = 20 on a swarm that fits one wave (scan_kernel<4,2,20>, one agent per warp), 12 closed-loop steps"""
    from multiagent_planning_b200 import scenarios
    N = 120
    pmin, pmax = scenarios.density_arena(N, density=2.0)
    po, pf = scenarios.random_test(N, pmin, pmax, 0.35, 2.0, seed=2020)
    P = dmpc.default_params(0, K=20)
    with dmpc.Solver(N, P, pmin=pmin, pmax=pmax, pf=pf) as s:
        l, pk, vk, ak = s.init_horizons(po)
        for _ in range(12):
            g, _ = _cmp_step(orc, P, s, pk, vk, ak, pf, l, pmin, pmax)
            l, pk, vk, ak = g["l_new"], g["p1"], g["v1"], g["a1"]


def test_outbound_and_infeasible_statuses(dmpc, orc, golden):
    """The failure statuses, produced for real and compared with the oracle.
    OUTBOUND (is_inbounds.m:1-6): with the reference's tolerance (+50 mm) the workspace rows of the QP imply
    it for an exact solver -- MATLAB only reaches it through quadprog's loosened ConstraintTolerance -- so the
    test tightens the parameter (inb_tol < 0) to make agents near a wall fail the test.
    INFEASIBLE: (a) a start outside the workspace (no slack to relax: reported at once, 0 retries);
    (b) the retry budget of solveSoftDMPCbound.m:102 exhausted (max_tries = 1 on a dense step)."""
    g = golden["kat_soft_bound"]
    args = (g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"], g["pmax"])
    P, s = _solver(dmpc, g, dmpc.SOFT_BOUND, inb_tol=-0.3)
    with s:
        out, o = _cmp_step(orc, P, s, *args)
        ob = (out["status"] & dmpc.ST_OUTBOUND) != 0
        assert ob.sum() >= 5 and ((out["status"][ob] & dmpc.ST_SOLVED) != 0).all()
        assert out["first_fail"] == int(np.nonzero(ob | ((out["status"] & 1) == 0))[0][0])
    P, s = _solver(dmpc, g, dmpc.SOFT_BOUND)
    pk = g["pk_prev"].copy()
    l = g["l"].copy(order="F")
    pk[0, 3] = g["pmax"][0] + 0.3
    l[0, :, 3] += pk[0, 3] - g["pk_prev"][0, 3]
    with s:
        out, _ = _cmp_step(orc, P, s, pk, g["vk_prev"], g["ak_prev"], g["pf"], l, g["pmin"], g["pmax"])
        assert out["status"][3] == dmpc.ST_INFEASIBLE and out["first_fail"] == 3
        assert np.array_equal(out["l_new"][:, :, 3], l[:, :, 3])          # the reference returns []: state kept
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config("C3")
    P = dmpc.default_params(0)
    with dmpc.Solver(500, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        s.init_horizons(cfg["po"])
        s.run(5)
        st = s.get_state()
    P1 = dmpc.default_params(0, max_tries=1)
    with dmpc.Solver(500, P1, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        out, o = _cmp_step(orc, P1, s, st["pk"], st["vk"], st["ak"], cfg["pf"], st["l"], cfg["pmin"], cfg["pmax"])
        inf = (out["status"] & dmpc.ST_INFEASIBLE) != 0
        assert inf.any() and (((out["status"][inf] >> 8) & 0xFF) == 1).all()


def _gloo_cuda_worker(rank, world, port, q, N, steps):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multiagent_planning_b200 import dmpc, scenarios, sharded
        pmin, pmax = scenarios.density_arena(N, density=1.5)
        po, pf = scenarios.random_test(N, pmin, pmax, 0.35, 2.0, seed=5)
        sh = sharded.ShardedDMPC(N, dmpc.default_params(0), pmin, pmax, po, pf)
        for _ in range(steps):
            sh.step()
        torch.cuda.synchronize()
        st = sh.gather_status()
        if rank == 0:
            q.put((sh.horizons(), st))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_two_ranks_on_one_gpu_equal_whole(dmpc):
    """The multi-GPU data path on the ONE GPU of the test box: two processes, each with a CudaBackend on
    cuda:0 solving its block with the real kernels, the per-step all-gather over gloo (NCCL refuses two ranks
    on one device).  Horizons and status words must be bit-equal to the single-handle run; N = 101 is not
    divisible by the world size (ragged last block)."""
    import torch.multiprocessing as mp
    from tests.test_sharded_gloo import _free_port
    from multiagent_planning_b200 import scenarios
    N, steps = 101, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_cuda_worker, args=(r, 2, port, q, N, steps)) for r in range(2)]
    for p in procs:
        p.start()
    l2, st2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pmin, pmax = scenarios.density_arena(N, density=1.5)
    po, pf = scenarios.random_test(N, pmin, pmax, 0.35, 2.0, seed=5)
    with dmpc.Solver(N, dmpc.default_params(0), pmin=pmin, pmax=pmax, pf=pf) as s:
        s.init_horizons(po)
        s.run(steps)
        st = s.get_state()
        assert np.array_equal(st["l"], l2) and np.array_equal(st["status"], st2)


def test_scenario_batch_equals_single_runs(dmpc):
    """n_scenarios independent swarms in ONE handle (failure_rate.m trial loops as a batch): trajectories, steps,
    goal / failure outcome of every scenario bit-equal to its own single-scenario run (which is oracle-gated
    above).  N = 50 is not a multiple of the 32-agent tile; the arenas differ per scenario; one scenario fails
    (two agents start inside each other) and stops at once while the others go on."""
    from multiagent_planning_b200 import scenarios
    N, S, steps = 50, 5, 130
    P = dmpc.default_params(0)
    scen = []
    for s in range(S):
        pmin, pmax = scenarios.density_arena(N, density=0.6 + 0.3 * s)
        po, pf = scenarios.random_test(N, pmin, pmax, 0.35, 2.0, seed=300 + s)
        scen.append((po, pf, pmin, pmax))
    scen[3][0][:, 7] = scen[3][0][:, 6] + np.array([0.05, 0.0, 0.0])       # a collision at k = 1 in scenario 3
    with dmpc.Solver(N, P, n_scenarios=S) as b:
        for s, c in enumerate(scen):
            b.set_scenario(s, *c)
        rb = b.run_batch(steps, stop_on_fail=True, record=True)
        last = [b.get_scenario(s) for s in range(S)]
        with pytest.raises(dmpc.DmpcError):
            b.run(3)                         # single-swarm entry point on a batched handle
        # again from the start, plain launches: same bits
        for s, c in enumerate(scen):
            b.set_scenario(s, *c)
        rb2 = b.run_batch(steps, stop_on_fail=True, mode=2, record=True)
    assert rb["first_fail_step"][3] == 0 and rb["first_fail_agent"][3] == 6 and rb["steps"][3] == 1
    assert rb["agent_steps"] == int(rb["steps"].sum()) * N
    for s, (po, pf, pmin, pmax) in enumerate(scen):
        with dmpc.Solver(N, P, pmin=pmin, pmax=pmax, pf=pf) as one:
            one.init_horizons(po)
            r = one.run(steps, stop_on_fail=True, record=True)
            st = one.get_state()
        assert r["steps"] == rb["steps"][s] == rb2["steps"][s] and r["reached"] == rb["reached"][s]
        assert r["first_fail_step"] == rb["first_fail_step"][s] and r["first_fail_agent"] == rb["first_fail_agent"][s]
        for k in ("pk", "vk", "ak"):
            assert np.array_equal(r[k], rb[k][s]), (s, k)
            assert np.array_equal(r[k], rb2[k][s]), (s, k)
        assert np.array_equal(st["l"], last[s]["l"]) and np.array_equal(st["pk"], last[s]["pk"])
        assert np.array_equal(st["status"], last[s]["status"])
    assert rb["reached"].sum() >= 3


def test_throughput_layout_vs_classic_layout(dmpc, orc, monkeypatch):
    """More than one wave of agents runs the two-role QP kernel (light agents with a 32-capacity active set,
    eight agents per SM; heavy agents routed to the 64-capacity workspaces by the scan kernel's prediction or
    on overflow).  The route an agent takes must not matter: (a) the layout is deterministic (two runs, same
    bits, although the queues are served in a timing-dependent order), (b) it agrees with the classic layout
    to rounding (the two kernels carry separately optimised copies of the solver) -- teacher-forced, so that
    rounding cannot be amplified by the closed loop --, status words equal, and (c) with the oracle.  A large
    dense swarm (N = 900: many heavy agents) and a batch of scenarios (8 x 100)."""
    from multiagent_planning_b200 import scenarios
    monkeypatch.setenv("DMPCB200_LAYOUT", "throughput")      # opt-in layout (the classic kernel is the default)
    N = 900
    pmin, pmax = scenarios.density_arena(N, density=1.5)
    po, pf = scenarios.random_test(N, pmin, pmax, 0.35, 2.0, seed=99)
    P = dmpc.default_params(0)
    a, c = scenarios.density_arena(100)
    scen = [scenarios.random_test(100, a, c, 0.35, 2.0, seed=700 + s) for s in range(8)]

    def batch():
        with dmpc.Solver(100, P, n_scenarios=8) as b:
            for i, (o, f) in enumerate(scen):
                b.set_scenario(i, o, f, a, c)
            return b.run_batch(40, stop_on_fail=False, record=True)

    with dmpc.Solver(N, P, pmin=pmin, pmax=pmax, pf=pf) as s:
        s.init_horizons(po)
        r1 = s.run(15, record=True, status_hist=True)
        states = []                                  # the inputs of every step of this run
        l, pk, vk, ak = s.init_horizons(po)
        for k in range(15):
            states.append((l, pk, vk, ak))
            g = s.step(pk, vk, ak, l)
            l, pk, vk, ak = g["l_new"], g["p1"], g["v1"], g["a1"]
            states[-1] += (g,)
        assert np.array_equal(r1["pk"][:, 15, :], pk)                    # (a) resident == host-stepped, odd step count
        s.init_horizons(po)
        r1b = s.run(15, record=True, status_hist=True)
        assert np.array_equal(r1["pk"], r1b["pk"]) and np.array_equal(r1["status_hist"], r1b["status_hist"])
        heavy = int((s.get_state()["diag"]["nact"] > 32).sum())
    b1, b1b = batch(), batch()
    for i in range(8):
        assert np.array_equal(b1["pk"][i], b1b["pk"][i])
    assert heavy >= 3                                                    # the 64-capacity route was really taken
    monkeypatch.setenv("DMPCB200_LAYOUT", "classic")
    with dmpc.Solver(N, P, pmin=pmin, pmax=pmax, pf=pf) as s:
        for k in (0, 3, 8, 14):
            l, pk, vk, ak, g = states[k]
            gc = s.step(pk, vk, ak, l)
            assert np.array_equal(gc["status"], g["status"])
            assert np.abs(gc["l_new"] - g["l_new"]).max() < 1e-9
        o = orc.step(oracle_params(orc, P), pk, vk, ak, pf, l, pmin, pmax)
        assert np.array_equal(g["status"] & 0xFFFF, o["status"] & 0xFFFF) and np.abs(g["l_new"] - o["l_new"]).max() < TOL
    b2 = batch()
    for i in range(8):
        assert b1["steps"][i] == b2["steps"][i] and np.abs(b1["pk"][i][:, :20] - b2["pk"][i][:, :20]).max() < 1e-9


def test_device_scenario_generation_and_monte_carlo_harness(dmpc, orc):
    """randomTest.m / randomExchange.m on the device (one CTA per scenario) == the oracle's restatement bit for
    bit; the failure_rate.m-shaped harness classifies a small batch consistently with single-scenario runs."""
    from multiagent_planning_b200 import montecarlo, scenarios
    N, S = 60, 7
    pmin, pmax = scenarios.density_arena(N)
    P = dmpc.default_params(0)
    for mode in (0, 1):
        with dmpc.Solver(N, P, n_scenarios=S) as b:
            po, pf = b.gen_scenarios(123, pmin, pmax, rmin_init=0.35, mode=mode)
            for s in range(S):
                opo, opf = orc.gen_scenario(123, s, mode, N, pmin, pmax, 0.35, P.c)
                assert np.array_equal(po[s], opo) and np.array_equal(pf[s], opf), (mode, s)
            st = b.get_scenario(2)
            assert np.array_equal(st["pk"], po[2]) and np.array_equal(st["l"][:, 0, :], po[2])    # initDMPC ran
        E = np.array([1.0, 1.0, 1.0 / P.c if mode == 0 else 1.0])[:, None, None]
        d = np.sqrt((((po[0][:, :, None] - po[0][:, None, :]) * E) ** 2).sum(0)) + 9 * np.eye(N)
        assert d.min() > 0.35 and (po[0] >= pmin[:, None]).all() and (po[0] <= pmax[:, None]).all()
        if mode == 1:
            assert sorted(map(tuple, po[1].T)) == sorted(map(tuple, pf[1].T)) and not (po[1] == pf[1]).all(0).any()
    r = montecarlo.failure_rate(N_vector=(12, 24), trials=6, seed=5)
    assert r["prob_dmpc"].shape == (2,) and ((r["success_dmpc"] <= r["feasible"]).all())
    assert (r["success_dmpc"] + r["failed_goal"] <= 1).all()
    ok = r["success_dmpc"][0] > 0
    assert ok.any() and np.isfinite(r["traj_time"][0][ok]).all() and (r["min_dist"][0][ok] >= 0.35 - 0.05).all()
    # trial 1 of the first size on its own: same steps, same post-processing figures
    pmin, pmax = scenarios.density_arena(12)
    po, pf = orc.gen_scenario(5, 1, 0, 12, pmin, pmax, 0.35, P.c)
    with dmpc.Solver(12, P, pmin=pmin, pmax=pmax, pf=pf) as s:
        s.init_horizons(po)
        one = s.run(149, stop_on_fail=2, record=True)
        assert one["steps"] == int(r["steps"][0, 1]) and bool(one["reached"]) == bool(r["feasible"][0, 1] and not r["failed_goal"][0, 1])
        if one["reached"]:
            pp = s.postprocess(one["pk"], one["vk"], one["ak"], want_interp=False)
            assert pp["traj_time"] == r["traj_time"][0, 1] and pp["violation"] == r["violation"][0, 1]


@pytest.mark.parametrize("name,layouts", [("C3", ("1", "7")), ("N2000", ("4", "9")), ("N2000", ("4", "11")), ("C4", ("5", "10"))])
def test_register_tile_scan_equals_per_agent_scan(dmpc, monkeypatch, name, layouts):
    """the two organisations of the neighbour scan (a warp per agent re-reading the tiles / a warp per tile with the
    neighbour's horizon in registers, looping over the CTA's agents) evaluate every pair with the same arithmetic:
    first violating step, neighbour sets and rows -- hence whole closed-loop steps -- must be bit-identical"""
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config(name)
    P = dmpc.default_params(cfg["variant"], **cfg["params"])
    res = []
    for lay in layouts:
        monkeypatch.setenv("DMPCB200_SCAN_LAYOUT", lay)
        with dmpc.Solver(cfg["N"], P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
            s.init_horizons(cfg["po"])
            r = s.run(10, record=True, status_hist=True)
            res.append((r, s.get_state()))
    (r1, s1), (r2, s2) = res
    assert np.array_equal(s1["diag"]["kstar"], s2["diag"]["kstar"]) and np.array_equal(s1["diag"]["nv"], s2["diag"]["nv"])
    assert (s1["diag"]["nv"] > 0).sum() > 50
    for k in ("pk", "vk", "ak", "status_hist"):
        assert np.array_equal(r1[k], r2[k]), k
    assert np.array_equal(s1["l"], s2["l"])


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from multiagent_planning_b200 import dmpc, scenarios, sharded
        cfg = scenarios.config("N100")
        P = dmpc.default_params(0)
        sh = sharded.ShardedDMPC(100, P, cfg["pmin"], cfg["pmax"], cfg["po"], cfg["pf"])
        for _ in range(10):
            sh.step()
        torch.cuda.synchronize()
        if rank == 0:
            q.put(sh.horizons())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpus_nccl_equal_one_gpu(dmpc):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from tests.test_sharded_gloo import _free_port
    from multiagent_planning_b200 import scenarios
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    l2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = scenarios.config("N100")
    with dmpc.Solver(100, dmpc.default_params(0), pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        s.init_horizons(cfg["po"])
        s.run(10)
        assert np.array_equal(s.get_state()["l"], l2)


def test_bound_step_and_host_timing(dmpc, orc, golden):
    """bind_step: pre-marshalled caller-owned buffers; same bits as step(); host phase timing is reported"""
    g = golden["kat_soft_bound"]
    P, s = _solver(dmpc, g, 0)
    f = lambda a: np.asfortranarray(np.array(a, dtype=np.float64))
    pk, vk, ak, l = f(g["pk_prev"]), f(g["vk_prev"]), f(g["ak_prev"]), f(g["l"])
    with s:
        ref = s.step(pk, vk, ak, l)
        out = dict(l_new=np.zeros_like(l), p1=np.zeros_like(pk), v1=np.zeros_like(pk), a1=np.zeros_like(pk),
                   status=np.zeros(200, np.int32), diag=np.zeros(200, dtype=ref["diag"].dtype))
        call = s.bind_step(pk, vk, ak, l, out)
        for _ in range(3):
            ff = call()
            assert ff == ref["first_fail"]
            for k in ("l_new", "p1", "v1", "a1", "status"):
                assert np.array_equal(out[k], ref[k]), k
        t = s.last_host_timing()
        assert set(t) == {"pack_us", "submit_us", "wait_us", "unpack_us"} and t["wait_us"] > 0
        with pytest.raises(dmpc.DmpcError):
            s.bind_step(pk.astype(np.float32), vk, ak, l, out)
        # pinned caller arrays: the kernel writes the outputs straight into them (no staging copy)
        import torch
        pin = lambda shape: torch.zeros(shape, dtype=torch.float64).pin_memory().numpy()
        out2 = dict(l_new=pin((200, 15, 3)).transpose(2, 1, 0), p1=pin((200, 3)).T, v1=pin((200, 3)).T,
                    a1=pin((200, 3)).T, status=np.zeros(200, np.int32), diag=np.zeros(200, dtype=ref["diag"].dtype))
        call2 = s.bind_step(pk, vk, ak, l, out2)
        for _ in range(2):
            assert call2() == ref["first_fail"]
            for k in ("l_new", "p1", "v1", "a1", "status"):
                assert np.array_equal(out2[k], ref[k]), k


def test_large_swarm_persistent_grid_vs_oracle(dmpc, orc):
    """N > 4 x (number of SMs): the QP kernel runs as a persistent grid with an agent queue and the scan
    kernel in its 8-agents-per-CTA layout with tile refill; teacher-forced against the oracle"""
    from multiagent_planning_b200 import scenarios
    N = 700
    pmin, pmax = scenarios.density_arena(N)
    po, pf = scenarios.random_test(N, pmin, pmax, 0.35, 2.0, seed=77)
    P = dmpc.default_params(0)
    with dmpc.Solver(N, P, pmin=pmin, pmax=pmax, pf=pf) as s:
        l, pk, vk, ak = s.init_horizons(po)
        for _ in range(3):
            g, _ = _cmp_step(orc, P, s, pk, vk, ak, pf, l, pmin, pmax)
            l, pk, vk, ak = g["l_new"], g["p1"], g["v1"], g["a1"]
        # and the resident loop gives the same bits as the host-stepped one
        s.init_horizons(po)
        r = s.run(3, record=True)
        assert np.array_equal(r["pk"][:, 3, :], pk)


def test_postprocess_vs_oracle_and_matlab(dmpc, orc, golden):
    """dmpcb200_postprocess (failure_rate.m:134-195) on the finished N=200 trial of failure_rate3.mat:
    against the oracle and against MATLAB's own workspace (time scaling bit-exact, splines 1e-12)"""
    from tests.conftest import raw_transition
    g = golden["postprocess_n200"]
    pk, vk, ak = raw_transition(g)
    N = pk.shape[2]
    P = dmpc.default_params(0, c=float(g["c"]), rmin=float(g["rmin"]), h=float(g["h"]))
    with dmpc.Solver(N, P, pf=g["pf"]) as s:
        r = s.postprocess(pk, vk, ak)
        o = orc.postprocess(pk, vk, ak, g["pf"], float(g["h"]), c=float(g["c"]), rmin=float(g["rmin"]))
        assert r["r_factor"] == o["r_factor"] == float(g["r_factor"]) and r["h_scaled"] == float(g["h_scaled"])
        assert r["nt"] == o["nt"] == int(g["nt"])
        for k in ("pk", "vk", "ak"):
            assert np.array_equal(r[k], o[k]), k                       # bit-exact time scaling
        sel = g["sel"]
        for k in ("p", "v", "a"):
            assert np.abs(r[k] - o[k]).max() < 1e-12, k
            assert np.abs(r[k][:, :, sel] - g[k]).max() < 1e-12, k     # MATLAB's spline
        assert np.array_equal(r["time_index"], g["time_index"]) and r["traj_time"] == float(g["traj_time"])
        assert r["violation"] == int(g["violation"]) == 1 and abs(r["min_dist"] - o["min_dist"]) < 1e-12
        assert abs(r["totdist"] - float(g["totdist"])) < 1e-8
        # figures only (no interpolated output requested)
        r2 = s.postprocess(pk, vk, ak, want_interp=False)
        assert r2["nt"] == r["nt"] and r2["min_dist"] == r["min_dist"] and r2["traj_time"] == r["traj_time"]
        with pytest.raises(dmpc.DmpcError):
            s.postprocess(pk[:, :3], vk[:, :3], ak[:, :3])            # too short for a not-a-knot spline


def test_postprocess_of_device_run(dmpc, orc):
    """whole pipeline: device-resident transition (N=100) -> post-processing, against the oracle's"""
    from multiagent_planning_b200 import scenarios
    cfg = scenarios.config("N100")
    P = dmpc.default_params(cfg["variant"])
    with dmpc.Solver(100, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        s.init_horizons(cfg["po"])
        run = s.run(149, record=True)
        assert run["reached"]
        r = s.postprocess(run["pk"], run["vk"], run["ak"])
        o = orc.postprocess(run["pk"], run["vk"], run["ak"], cfg["pf"], P.h, c=P.c, rmin=P.rmin)
        assert r["r_factor"] == o["r_factor"] and r["nt"] == o["nt"] and r["violation"] == o["violation"]
        assert np.array_equal(r["pk"], o["pk"]) and np.abs(r["p"] - o["p"]).max() < 1e-12
        assert np.array_equal(r["time_index"], o["time_index"]) and abs(r["totdist"] - o["totdist"]) < 1e-8
        assert abs(r["min_dist"] - o["min_dist"]) < 1e-12
