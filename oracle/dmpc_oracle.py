"""ctypes binding + numpy cross-checks for the CPU ORACLE (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (multiagent_planning_b200) never does.

Besides the binding to oracle/liboracle.so (C restatement of dmpc/matlab, see dmpc_oracle.c)
this file holds
  * numpy restatements of the small helpers (getPosMat.m, getDeltaMat.m, ...) used to cross
    check the C code, and
  * qp_pdip(): an independent dense Mehrotra primal-dual interior-point QP solver used ONLY to
    cross-check the Goldfarb-Idnani solver of the oracle on the same dense QPs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

VARIANT_SOFT_BOUND, VARIANT_SOFT_BOUND2, VARIANT_HARD, VARIANT_HARD_ONDEMAND = 0, 1, 2, 3
ST_SOLVED, ST_COLL, ST_INFEASIBLE, ST_OUTBOUND, ST_QPFAIL = 1, 2, 4, 8, 16


class Params(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("variant", C.c_int32), ("max_tries", C.c_int32), ("neigh_mode", C.c_int32),
        ("h", C.c_double), ("rmin", C.c_double), ("c", C.c_double), ("alim", C.c_double),
        ("Q1", C.c_double), ("S1", C.c_double), ("term", C.c_double),
        ("Q_far", C.c_double), ("Q_near", C.c_double), ("S_free", C.c_double),
        ("near_radius", C.c_double), ("slack_lb", C.c_double), ("neigh_factor", C.c_double),
        ("coll_tol", C.c_double), ("inb_tol", C.c_double), ("hard_radius", C.c_double),
        ("init_div", C.c_double),
    ]


class Diag(C.Structure):
    _fields_ = [
        ("kstar", C.c_int32), ("nv", C.c_int32), ("tries", C.c_int32), ("qp_iters", C.c_int32),
        ("kkt_stat", C.c_double), ("kkt_prim", C.c_double), ("kkt_comp", C.c_double),
        ("kkt_dual", C.c_double), ("min_dist", C.c_double), ("objective", C.c_double),
        ("n_act_box", C.c_int32), ("n_act_pos", C.c_int32), ("n_act_row", C.c_int32), ("n_act_eps", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with gcc (oracle/Makefile)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "dmpc_oracle.c")
    hdr = os.path.join(_HERE, "dmpc_oracle.h")
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip, u8p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        L.orc_default_params.argtypes = [C.POINTER(Params), C.c_int]
        L.orc_model_mats.argtypes = [C.c_double, C.c_int, dp, dp, dp, dp]
        L.orc_init_dmpc.argtypes = [dp, dp, C.c_double, C.c_int, C.c_double, dp, dp, dp]
        L.orc_check_coll.argtypes = [C.POINTER(Params), dp, dp, C.c_int, C.c_int, C.c_int, u8p, u8p, dp]
        L.orc_check_coll.restype = C.c_int
        L.orc_coll_constr.argtypes = [C.POINTER(Params), dp, dp, dp, C.c_int, C.c_int, dp, C.c_int,
                                      u8p, dp, C.c_int, dp, dp, ip]
        L.orc_coll_constr.restype = C.c_int
        L.orc_solve_agent.argtypes = [C.POINTER(Params), dp, dp, dp, dp, C.c_int, dp, C.c_int, dp, dp,
                                      dp, dp, dp, C.POINTER(Diag)]
        L.orc_solve_agent.restype = C.c_int
        L.orc_step.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp,
                               dp, dp, dp, dp, ip, C.POINTER(Diag), C.c_int]
        L.orc_step.restype = C.c_int
        L.orc_reached_goal.argtypes = [dp, dp, C.c_int, C.c_double, dp]
        L.orc_reached_goal.restype = C.c_int
        L.orc_is_inbounds.argtypes = [dp, dp, dp, C.c_double]
        L.orc_is_inbounds.restype = C.c_int
        L.orc_prop_state.argtypes = [C.c_double, C.c_int, dp, dp, dp, dp, dp]
        L.orc_qp_gi.argtypes = [C.c_int, dp, dp, C.c_int, dp, C.c_int, dp, dp, dp, dp, dp, ip,
                                C.POINTER(Diag)]
        L.orc_qp_gi.restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f(a):
    """column-major contiguous float64 copy (MATLAB layout)."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def default_params(variant: int = VARIANT_SOFT_BOUND, **kw) -> Params:
    P = Params()
    lib().orc_default_params(C.byref(P), variant)
    for k, v in kw.items():
        if not hasattr(P, k):
            raise AttributeError(k)
        setattr(P, k, v)
    return P


def model_mats(h: float, K: int):
    n = 3 * K
    A_p, A_v, Delta = (np.zeros((n, n), order="F") for _ in range(3))
    A0 = np.zeros((n, 6), order="F")
    lib().orc_model_mats(h, K, _dp(A_p), _dp(A_v), _dp(A0), _dp(Delta))
    return A_p, A_v, A0, Delta


def init_dmpc(po, pf, h, K, init_div=10.0):
    po, pf = _f(po).ravel(), _f(pf).ravel()
    p, v, a = (np.zeros((3, K), order="F") for _ in range(3))
    lib().orc_init_dmpc(_dp(po), _dp(pf), h, K, init_div, _dp(p), _dp(v), _dp(a))
    return p, v, a


def check_coll(P: Params, p3, l, n: int, k: int):
    """CheckCollSoftDMPC: n 0-based, k 1-based. l is (3,K,N)."""
    l = _f(l)
    N = l.shape[2]
    p3 = _f(p3).ravel()
    viol = np.zeros(N, np.uint8)
    vc = np.zeros(N, np.uint8)
    md = C.c_double()
    u8p = C.POINTER(C.c_uint8)
    any_ = lib().orc_check_coll(C.byref(P), _dp(p3), _dp(l), N, n, k, viol.ctypes.data_as(u8p),
                                vc.ctypes.data_as(u8p), C.byref(md))
    return bool(any_), viol.astype(bool), vc.astype(bool), md.value


def coll_constr(P: Params, p3, po, vo, n: int, k: int, l, mask=None):
    l = _f(l)
    K, N = l.shape[1], l.shape[2]
    cap = max(N - 1, 1)
    Ain = np.zeros((cap, 3 * K), order="F")
    bin_ = np.zeros(cap)
    pd = np.zeros(cap)
    idx = np.zeros(cap, np.int32)
    m8 = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    u8p = C.POINTER(C.c_uint8)
    nr = lib().orc_coll_constr(C.byref(P), _dp(_f(p3).ravel()), _dp(_f(po).ravel()), _dp(_f(vo).ravel()),
                               n, k, _dp(l), N, None if m8 is None else m8.ctypes.data_as(u8p),
                               _dp(Ain), cap, _dp(bin_), _dp(pd), idx.ctypes.data_as(C.POINTER(C.c_int32)))
    return Ain[:nr].copy(), bin_[:nr].copy(), pd[:nr].copy(), idx[:nr].copy()


def solve_agent(P: Params, po, pf, vo, ao, n: int, l, pmin, pmax):
    """Returns (status, p, v, a, diag) with p,v,a (3,K) arrays (garbage unless status&ST_SOLVED)."""
    l = _f(l)
    K, N = l.shape[1], l.shape[2]
    p, v, a = (np.zeros((3, K), order="F") for _ in range(3))
    d = Diag()
    st = lib().orc_solve_agent(C.byref(P), _dp(_f(po).ravel()), _dp(_f(pf).ravel()), _dp(_f(vo).ravel()),
                               _dp(_f(ao).ravel()), n, _dp(l), N, _dp(_f(pmin).ravel()),
                               _dp(_f(pmax).ravel()), _dp(p), _dp(v), _dp(a), C.byref(d))
    return st, p, v, a, d


def step(P: Params, pk, vk, ak, pf, l_prev, pmin, pmax, n0=0, n1=None, nthreads=1, want_diag=False):
    """One Jacobi MPC step (agents n0..n1).  pk,vk,ak,pf: (3,N); l_prev (3,K,N)."""
    l_prev = _f(l_prev)
    K, N = l_prev.shape[1], l_prev.shape[2]
    n1 = N if n1 is None else n1
    pk, vk, ak, pf = _f(pk), _f(vk), _f(ak), _f(pf)
    l_new = l_prev.copy(order="F")
    p1, v1, a1 = pk.copy(order="F"), vk.copy(order="F"), ak.copy(order="F")
    status = np.zeros(N, np.int32)
    diags = (Diag * N)() if want_diag else None
    ff = lib().orc_step(C.byref(P), N, n0, n1, _dp(pk), _dp(vk), _dp(ak), _dp(pf), _dp(l_prev),
                        _dp(_f(pmin).ravel()), _dp(_f(pmax).ravel()), _dp(l_new), _dp(p1), _dp(v1), _dp(a1),
                        status.ctypes.data_as(C.POINTER(C.c_int32)), diags, nthreads)
    out = dict(l_new=l_new, p1=p1, v1=v1, a1=a1, status=status, first_fail=ff)
    if want_diag:
        out["diag"] = diags
    return out


def reached_goal(p, pf, tol):
    p, pf = _f(p), _f(pf)
    md = C.c_double()
    r = lib().orc_reached_goal(_dp(p), _dp(pf), p.shape[1], tol, C.byref(md))
    return bool(r), md.value


def qp_gi(H, f, A, b, lb, ub):
    """Exact dense QP (oracle's solver). Returns (rc, x, lam, iters, diag)."""
    H = _f(H)
    n = H.shape[0]
    A = _f(np.zeros((0, n)) if A is None else A).reshape(-1, n, order="F")
    m = A.shape[0]
    A = np.asfortranarray(A)
    b = _f(np.zeros(0) if b is None else b).ravel()
    x = np.zeros(n)
    lam = np.zeros(m + 2 * n)
    it = C.c_int32()
    d = Diag()
    rc = lib().orc_qp_gi(n, _dp(H), _dp(_f(f).ravel()), m, _dp(A), max(m, 1), _dp(b), _dp(_f(lb).ravel()),
                         _dp(_f(ub).ravel()), _dp(x), _dp(lam), C.byref(it), C.byref(d))
    return rc, x, lam, it.value, d


# ---------------------------------------------------------------------------------------------
# numpy restatements (cross-checks of the C code; small sizes only)
# ---------------------------------------------------------------------------------------------

def np_model_mats(h: float, K: int):
    """getPosMat.m:1-23 / dmpc_soft_bound.m:81-108 / getDeltaMat.m:1-9 with dense matrices,
    evaluated with the same recurrence order as the reference."""
    Aux = np.block([[np.eye(3), h * np.eye(3)], [np.zeros((3, 3)), np.eye(3)]])
    b = np.vstack([h ** 2 / 2 * np.eye(3), h * np.eye(3)])
    prev = np.zeros((6, 3 * K))
    A_p, A_v, A0 = [], [], []
    Ai = np.eye(6)
    for k in range(K):
        add = np.zeros((6, 3 * K))
        add[:, 3 * k:3 * k + 3] = b
        new = np.zeros_like(prev)
        # Aux*prev_row + add_b, written per block so that no FMA/accumulation-order ambiguity exists
        new[:3] = prev[:3] + h * prev[3:]
        new[3:] = prev[3:]
        new = new + add
        A_p.append(new[:3])
        A_v.append(new[3:])
        prev = new
        Ai = np.vstack([Ai[:3] + h * Ai[3:], Ai[3:]])
        A0.append(Ai[:3])
    n = 3 * K
    Delta = np.eye(n)
    for i in range(3, n):
        Delta[i, i - 3] = -1.0
    return np.vstack(A_p), np.vstack(A_v), np.vstack(A0), Delta


def np_dense_qp(P: Params, po, pf, vo, ao, n, l, pmin, pmax, slack_lb=None, term=None):
    """Dense (H,f,A,b,lb,ub) of solveSoftDMPCbound.m:60-98 built with numpy matrix products,
    literally as the reference writes them.  Soft variants only.  Returns None when no QP is
    formed (k==1 collision)."""
    l = np.asarray(l, dtype=np.float64)
    K, N = l.shape[1], l.shape[2]
    A, A_v, A0, Delta = np_model_mats(P.h, K)
    E1 = np.diag([1, 1, 1 / P.c])
    E2 = np.diag([1, 1, 1 / P.c ** 2])
    po, pf, vo, ao = (np.asarray(t, float).ravel() for t in (po, pf, vo, ao))
    prev_p = l[:, :, n]
    rows, bins, dists = [], [], []
    for k in range(1, K + 1):
        d = np.array([np.linalg.norm(E1 @ (prev_p[:, k - 1] - l[:, k - 1, i])) if i != n else np.inf
                      for i in range(N)])
        if (d < P.rmin).any():
            if k == 1 and d.min() < P.rmin - P.coll_tol:
                return None
            if P.variant == VARIANT_SOFT_BOUND2 and k == 1:
                continue
            k_ctr = k - 1 if P.variant == VARIANT_SOFT_BOUND2 else k
            for i in np.nonzero(d < P.neigh_factor * P.rmin)[0]:
                p = prev_p[:, k - 1]
                diff = E2 @ (p - l[:, k - 1, i])
                r = d[i] * (P.rmin - d[i] + diff @ p / d[i]) - diff @ A0[3 * (k_ctr - 1):3 * k_ctr] @ np.r_[po, vo]
                dm = np.zeros(3 * K)
                dm[3 * (k_ctr - 1):3 * k_ctr] = diff
                rows.append(-dm @ A)
                bins.append(-r)
                dists.append(d[i])
            break
    nv = len(rows)
    if nv == 0 and np.linalg.norm(po - pf) >= P.near_radius:
        q, s = P.Q_far, P.S_free
    elif nv == 0:
        q, s = P.Q_near, P.S_free
    else:
        q, s = P.Q1, P.S1
    n3 = 3 * K
    Q = np.zeros((n3, n3))
    Q[-3:, -3:] = q * np.eye(3)
    R = np.eye(n3)
    S = s * np.eye(n3)
    x0 = np.r_[po, vo]
    H = np.zeros((n3 + nv, n3 + nv))
    H[:n3, :n3] = 2 * (A.T @ Q @ A + Delta.T @ S @ Delta + R)
    H[n3:, n3:] = 2 * np.eye(nv)
    ao_1 = np.r_[ao, np.zeros(n3 - 3)]
    f = np.zeros(n3 + nv)
    f[:n3] = -2 * (np.tile(pf, K) @ Q @ A - (A0 @ x0) @ Q @ A + ao_1 @ S @ Delta)
    f[n3:] = P.term if term is None else term
    Ain = np.zeros((nv + 2 * n3, n3 + nv))
    if nv:
        Ain[:nv, :n3] = np.array(rows)
        Ain[:nv, n3:] = np.diag(dists)
    Ain[nv:nv + n3, :n3] = A
    Ain[nv + n3:, :n3] = -A
    bin_ = np.r_[np.array(bins), np.tile(np.asarray(pmax, float).ravel(), K) - A0 @ x0,
                 -np.tile(np.asarray(pmin, float).ravel(), K) + A0 @ x0]
    lb = np.r_[-P.alim * np.ones(n3), (P.slack_lb if slack_lb is None else slack_lb) * np.ones(nv)]
    ub = np.r_[P.alim * np.ones(n3), np.zeros(nv)]
    return H, f, Ain, bin_, lb, ub


def qp_pdip(H, f, A, b, lb, ub, tol=1e-11, max_iter=200):
    """Independent dense Mehrotra predictor-corrector IPM:  min 1/2 x'Hx + f'x, Ax<=b, lb<=x<=ub.
    Cross-check only (numpy, slow).  Returns (x, info)."""
    H = np.asarray(H, float)
    n = H.shape[0]
    G = np.vstack([A, np.eye(n), -np.eye(n)]) if A is not None and len(A) else np.vstack([np.eye(n), -np.eye(n)])
    h = np.r_[b if A is not None and len(A) else np.zeros(0), ub, -np.asarray(lb)]
    m = G.shape[0]
    scale = max(1.0, np.abs(f).max())
    x = np.clip(np.linalg.solve(H, -f), np.asarray(lb) * 0.5, np.asarray(ub) * 0.5) if n else np.zeros(0)
    s = np.maximum(h - G @ x, 1e-2)
    z = np.ones(m) * scale * 1e-3
    for it in range(max_iter):
        rd = H @ x + f + G.T @ z
        rp = G @ x + s - h
        mu = s @ z / m
        if max(np.abs(rd).max() / scale, np.abs(rp).max(), mu / scale) < tol:
            return x, dict(iters=it, converged=True, mu=mu, z=z)
        W = z / s
        M = H + G.T @ (W[:, None] * G)
        Lc = np.linalg.cholesky(M)

        def solve(rc):
            rhs = -rd + G.T @ (rc / s - W * rp)
            dx = np.linalg.solve(Lc.T, np.linalg.solve(Lc, rhs))
            ds = -rp - G @ dx
            dz = -(rc + z * ds) / s
            return dx, ds, dz

        def steplen(v, dv):
            neg = dv < 0
            return min(1.0, (-v[neg] / dv[neg]).min()) if neg.any() else 1.0

        dxa, dsa, dza = solve(s * z)
        aa = min(steplen(s, dsa), steplen(z, dza))
        mua = (s + aa * dsa) @ (z + aa * dza) / m
        sigma = (mua / mu) ** 3
        dx, ds, dz = solve(s * z + dsa * dza - sigma * mu)
        al = min(1.0, 0.995 * min(steplen(s, ds), steplen(z, dz)))
        x, s, z = x + al * dx, s + al * ds, z + al * dz
    return x, dict(iters=max_iter, converged=False, mu=mu, z=z)


# ---------------------------------------------------------------------------------------------
# scenario generation + closed loop (oracle side; the product has its own in scenarios.py/driver)
# ---------------------------------------------------------------------------------------------

def random_test(N, pmin, pmax, rmin, c, rng, max_iter=200000):
    """randomTest.m:1-57 restated: rejection sampling of N points with pairwise ellipsoid
    distance > rmin, for start and goal sets.  MATLAB's RNG stream cannot be reproduced, so the
    generator is numpy's; returns po, pf as (3,N)."""
    pmin, pmax = np.asarray(pmin, float), np.asarray(pmax, float)
    E1 = np.array([1.0, 1.0, 1.0 / c])

    def sample_set():
        while True:
            pts = np.zeros((3, N))
            pts[:, 0] = pmin + (pmax - pmin) * rng.random(3)
            ok = True
            for n in range(1, N):
                tries = 0
                while True:
                    cand = pmin + (pmax - pmin) * rng.random(3)
                    dist = np.sqrt((((pts[:, :n] - cand[:, None]) * E1[:, None]) ** 2).sum(0))
                    tries += 1
                    if (dist > rmin).all():
                        pts[:, n] = cand
                        break
                    if tries > max_iter:
                        ok = False
                        break
                if not ok:
                    break
            if ok:
                return pts

    return sample_set(), sample_set()


def simulate(P: Params, po, pf, pmin, pmax, max_steps, tol=0.01, nthreads=1, record=False, stop_on_fail=True):
    """Closed loop of test/failure_rate.m:99-127: step 1 = initDMPC for all agents, steps k>1 = the
    per-agent solve against l of step k-1.  Returns dict with pk,vk,ak (3,steps,N), final l,
    per-step status, reached_goal, steps."""
    po, pf = _f(po), _f(pf)
    N, K = po.shape[1], P.K
    l = np.zeros((3, K, N), order="F")
    for n in range(N):
        p, v, a = init_dmpc(po[:, n], pf[:, n], P.h, K, P.init_div)
        l[:, :, n] = p
    pk = [l[:, 0, :].copy()]
    vk = [np.zeros((3, N))]
    ak = [np.zeros((3, N))]
    hist = []
    reached, k, failed = False, 1, -1
    reached, _ = reached_goal(pk[-1], pf, tol)
    while not reached and k < max_steps:
        if record:
            hist.append(dict(l=l.copy(), pk=pk[-1].copy(), vk=vk[-1].copy(), ak=ak[-1].copy()))
        out = step(P, pk[-1], vk[-1], ak[-1], pf, l, pmin, pmax, nthreads=nthreads, want_diag=record)
        if record:
            hist[-1]["out"] = out
        if out["first_fail"] >= 0:
            failed = out["first_fail"]
            if stop_on_fail:
                break
        l = out["l_new"]
        pk.append(out["p1"]); vk.append(out["v1"]); ak.append(out["a1"])
        reached, _ = reached_goal(pk[-1], pf, tol)
        k += 1
    return dict(pk=np.stack(pk, 1), vk=np.stack(vk, 1), ak=np.stack(ak, 1), l=l, steps=k,
                reached_goal=reached, failed_agent=failed, hist=hist)


# ---- post-processing of a finished transition (test/failure_rate.m:134-195) ----------------------------
def spline_not_a_knot(x, y, xq):
    """MATLAB `spline(x, y, xq)` for row-wise data y (..., n), n >= 4: cubic spline with not-a-knot end
    conditions in the slope (Hermite) form of spline.m, evaluated like ppval (right-continuous pieces,
    the last point in the last piece).  Pure numpy."""
    x = np.asarray(x, float)
    y = np.asarray(y, float)
    n = x.size
    dx = np.diff(x)
    dd = np.diff(y, axis=-1) / dx
    A = np.zeros((n, n))
    b = np.zeros(y.shape)
    for i in range(1, n - 1):
        A[i, i - 1], A[i, i], A[i, i + 1] = dx[i], 2.0 * (dx[i - 1] + dx[i]), dx[i - 1]
        b[..., i] = 3.0 * (dx[i] * dd[..., i - 1] + dx[i - 1] * dd[..., i])
    x31, xn = x[2] - x[0], x[-1] - x[-3]
    A[0, 0], A[0, 1] = dx[1], x31
    b[..., 0] = ((dx[0] + 2.0 * x31) * dx[1] * dd[..., 0] + dx[0] ** 2 * dd[..., 1]) / x31
    A[-1, -2], A[-1, -1] = xn, dx[-2]
    b[..., -1] = (dx[-1] ** 2 * dd[..., -2] + (2.0 * xn + dx[-1]) * dx[-2] * dd[..., -1]) / xn
    s = np.linalg.solve(A, b.reshape(-1, n).T).T.reshape(y.shape)
    xq = np.asarray(xq, float)
    idx = np.clip(np.searchsorted(x, xq, side="right") - 1, 0, n - 2)
    t = xq - x[idx]
    hx = dx[idx]
    y0, y1, s0, s1, d = y[..., idx], y[..., idx + 1], s[..., idx], s[..., idx + 1], dd[..., idx]
    c2 = (3.0 * d - 2.0 * s0 - s1) / hx
    c3 = (s0 - 2.0 * d + s1) / hx ** 2
    return y0 + t * (s0 + t * (c2 + t * c3))


def gen_scenario(seed, scen, mode, N, pmin, pmax, rmin, c):
    """randomTest.m (mode 0) / randomExchange.m (mode 1), scenario `scen` of a batch -> po, pf (3, N)"""
    po, pf = np.zeros((3, N), order="F"), np.zeros((3, N), order="F")
    f = lib().orc_gen_scenario
    f.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                  C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f.restype = None
    f(int(seed), int(scen), int(mode), int(N), _dp(_f(pmin).ravel()), _dp(_f(pmax).ravel()), float(rmin), float(c),
      _dp(po), _dp(pf))
    return po, pf


def postprocess(pk, vk, ak, pf, h, c=2.0, rmin=0.35, vmax=2.0, amax=1.0, Ts=0.01, coll_tol=0.05,
                goal_radius=0.05, want_interp=True):
    """failure_rate.m:134-195 for one finished transition.  pk, vk, ak: (3, S, N) as the loop left them
    (column k = state after MPC step k).  Returns the time-scaled pk, vk, ak, the 100 Hz interpolation
    p, v, a (3, nt, N), and the trial's figures: r_factor, h_scaled, T, violation (post-interpolation
    collision check), min_dist, totdist, time_index (N), traj_time."""
    pk, vk, ak = (np.array(x, dtype=np.float64, order="F") for x in (pk, vk, ak))
    S, N = pk.shape[1], pk.shape[2]
    pf = np.asarray(pf, float).reshape(3, N, order="F")
    with np.errstate(divide="ignore"):
        ak_mod = amax / np.sqrt((ak[0] ** 2 + ak[1] ** 2) + ak[2] ** 2)       # :141
        vk_mod = vmax / np.sqrt((vk[0] ** 2 + vk[1] ** 2) + vk[2] ** 2)       # :142
    r_factor = min(ak_mod.min(), vk_mod.min())                                 # :144
    h_scaled = h / np.sqrt(r_factor)                                           # :145
    T = (S - 1) * h_scaled          # :148  (k - 2) h_scaled with k = S + 1 when the loop ended
    tk = np.arange(S) * h_scaled    # :149  0:h_scaled:T
    nt = int(np.floor(T / Ts + 1e-9)) + 1
    t = np.arange(nt) * Ts          # :151  0:Ts:T
    hh = h_scaled ** 2 / 2
    for k in range(S - 1):          # :156-162
        ak[:, k, :] = ak[:, k, :] * r_factor
        vk[:, k + 1, :] = vk[:, k, :] + h_scaled * ak[:, k, :]
        pk[:, k + 1, :] = (pk[:, k, :] + h_scaled * vk[:, k, :]) + hh * ak[:, k, :]
    out = dict(pk=pk, vk=vk, ak=ak, r_factor=r_factor, h_scaled=h_scaled, T=T, tk=tk, t=t, nt=nt)
    p = spline_not_a_knot(tk, np.moveaxis(pk, 1, -1), t)      # (3, N, nt)   :164-168
    if want_interp:
        out["p"] = np.moveaxis(p, -1, 1)
        out["v"] = np.moveaxis(spline_not_a_knot(tk, np.moveaxis(vk, 1, -1), t), -1, 1)
        out["a"] = np.moveaxis(spline_not_a_knot(tk, np.moveaxis(ak, 1, -1), t), -1, 1)
    # :170-181 pairwise check on the interpolated positions, E1 = diag(1, 1, 1/c)
    md = np.inf
    for i in range(N - 1):
        d = p[:, i + 1:, :] - p[:, i:i + 1, :]
        dist = np.sqrt((d[0] ** 2 + d[1] ** 2) + (d[2] / c) ** 2)
        md = min(md, dist.min())
    out["min_dist"] = md
    out["violation"] = int(md < rmin - coll_tol)
    dp = np.diff(p, axis=-1)
    out["totdist"] = float(np.sqrt((dp[0] ** 2 + dp[1] ** 2) + dp[2] ** 2).sum())   # :183
    dg = np.sqrt(((p - pf[:, :, None]) ** 2).sum(0))                           # :185-193
    far = dg >= goal_radius
    last = np.where(far.any(1), nt - 1 - np.argmax(far[:, ::-1], axis=1), -1)
    out["time_index"] = np.where(last >= 0, last + 2, 0).astype(np.int32)      # 1-based hola + 1
    out["traj_time"] = float(out["time_index"].max() * Ts)
    return out
