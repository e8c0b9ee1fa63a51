"""Host-side mirror of the reference's DMPC entry surface over the C-ABI of libdmpc_b200.so.

Two layers:

* ``Solver`` -- thin object over one library handle (one GPU, agents [n0, n1) of N): batched
  ``step`` (the body of ``for n = 1:N`` in test/failure_rate.m:100-119), device-resident ``run``
  (the whole ``while ~reached_goal`` loop), and the per-agent helper drop-ins.
* functions with the reference's MATLAB names and argument order (dmpc/matlab/*.m):
  ``solveSoftDMPCbound``, ``solveSoftDMPCbound2``, ``solveHardDMPC``, ``solveHardDMPCOnDemand``,
  ``CheckCollSoftDMPC``, ``CollConstrSoftDMPC``, ``CollConstrSoftDMPC2``, ``CollConstrHardDMPC``,
  ``CollConstrHardDMPCOnDemand``, ``propStatedmpc``, ``getPosMat``, ``getDeltaMat``, ``initDMPC``,
  ``is_inbounds``, ``ReachedGoal`` and the dec-iSCP-named aliases ``CollConstr`` / ``propState``
  that BASELINE.json lists.  Agent / horizon indices are 1-based in these functions exactly like
  MATLAB; arrays are numpy, column-major semantics (l is (3, K, N)).

All compute runs in hand-written sm_100a CUDA; nothing here computes on the CPU except trivial
argument marshalling and the closed-form model matrices (host code in the reference too).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Diag, DmpcError, Params, Post

SOFT_BOUND, SOFT_BOUND2, HARD, HARD_ONDEMAND = 0, 1, 2, 3
ST_SOLVED, ST_COLL, ST_INFEASIBLE, ST_OUTBOUND, ST_QPFAIL, ST_OVERFLOW = 1, 2, 4, 8, 16, 32

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


def _f(a, shape=None):
    if (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.f_contiguous
            and (shape is None or a.shape == shape)):
        return a  # already in the library's layout: no copy, no new object
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = np.asfortranarray(a.reshape(shape, order="F"))
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def default_params(variant: int = SOFT_BOUND, **kw) -> Params:
    """Reference defaults (test/failure_rate.m:7-27, solveSoftDMPCbound.m) with overrides."""
    P = Params()
    _lib.lib().dmpcb200_default_params(C.byref(P), int(variant))
    for k, v in kw.items():
        if not hasattr(P, k):
            raise TypeError(f"unknown parameter {k}")
        setattr(P, k, v)
    return P


def device_count() -> int:
    return int(_lib.lib().dmpcb200_device_count())


class Solver:
    """One library handle: N agents, this handle solves agents [n0, n1) on CUDA device `device`."""

    def __init__(self, N: int, params: Params | None = None, n0: int = 0, n1: int | None = None,
                 device: int = 0, max_rows: int = 0, pmin=None, pmax=None, pf=None, n_scenarios: int = 1):
        self.L = _lib.lib()
        self.P = params if params is not None else default_params()
        self.N, self.K = int(N), int(self.P.K)
        self.n0, self.n1 = int(n0), int(N if n1 is None else n1)
        self.device = int(device)
        self.S = int(n_scenarios)
        h = C.c_void_p()
        _lib.check(self.L.dmpcb200_create(C.byref(self.P), self.N, self.n0, self.n1, self.S, self.device,
                                          int(max_rows), C.byref(h)), "dmpcb200_create")
        self.h = h
        if pmin is not None:
            self.set_bounds(pmin, pmax)
        if pf is not None:
            self.set_goals(pf)

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.dmpcb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- scenario ---------------------------------------------------------------------------------
    def set_bounds(self, pmin, pmax):
        pmin, pmax = _f(pmin).ravel(), _f(pmax).ravel()
        _lib.check(self.L.dmpcb200_set_bounds(self.h, _p(pmin), _p(pmax)), "set_bounds")

    def set_goals(self, pf):
        pf = _f(pf, (3, self.N))
        _lib.check(self.L.dmpcb200_set_goals(self.h, _p(pf)), "set_goals")

    def set_static_obstacles(self, n_cmd):
        """agents [n_cmd, N) are static obstacles (the C++ port's N_cmd < N, dmpc.cpp:1633-1649)"""
        _lib.check(self.L.dmpcb200_set_static_obstacles(self.h, int(n_cmd)), "set_static_obstacles")
        self.n1 = int(n_cmd)

    def init_horizons(self, po):
        """initDMPC.m for all agents.  Returns l (3,K,N), p1, v1, a1 (3,N)."""
        po = _f(po, (3, self.N))
        l = np.zeros((3, self.K, self.N), order="F")
        p1, v1, a1 = (np.zeros((3, self.N), order="F") for _ in range(3))
        _lib.check(self.L.dmpcb200_init_horizons(self.h, _p(po), _p(l), _p(p1), _p(v1), _p(a1)), "init_horizons")
        return l, p1, v1, a1

    # -- scenario batching (test/failure_rate.m trial loops as one batch) ------------------------------
    def set_scenario(self, s, po, pf, pmin, pmax):
        """start points, goals (3,N) and workspace box of scenario s; runs initDMPC.m for its agents"""
        po, pf = _f(po, (3, self.N)), _f(pf, (3, self.N))
        pmin, pmax = _f(pmin).ravel(), _f(pmax).ravel()
        _lib.check(self.L.dmpcb200_set_scenario(self.h, int(s), _p(po), _p(pf), _p(pmin), _p(pmax)), "set_scenario")

    def gen_scenarios(self, seed, pmin, pmax, rmin_init=None, mode=0, want_points=True):
        """randomTest.m (mode 0) / randomExchange.m (mode 1) for every scenario of the handle, on the device;
        sets bounds and goals and runs initDMPC.m.  Returns po, pf of shape (S, 3, N) when want_points."""
        pmin, pmax = _f(pmin).ravel(), _f(pmax).ravel()
        po = pf = None
        if want_points:
            po, pf = np.zeros(3 * self.N * self.S), np.zeros(3 * self.N * self.S)
        _lib.check(self.L.dmpcb200_gen_scenarios(self.h, int(seed), int(mode),
                                                 float(self.P.rmin if rmin_init is None else rmin_init), _p(pmin),
                                                 _p(pmax), _p(po), _p(pf)), "gen_scenarios")
        if want_points:
            shp = lambda a: np.stack([np.asfortranarray(a[3 * self.N * s:3 * self.N * (s + 1)].reshape((3, self.N), order="F"))
                                      for s in range(self.S)])
            return shp(po), shp(pf)

    def run_batch(self, max_steps, stop_on_fail=True, mode=0, record=False):
        """the closed loop of every scenario (failure_rate.m:99-133 per trial).  Returns per-scenario arrays
        steps, reached, first_fail_step, first_fail_agent, goal_dist (+ pk, vk, ak of shape (S, 3, T, N) in
        column-major blocks when record) and the device time / agent-steps of the call."""
        S, N = self.S, self.N
        steps, reached, fs, fa = (np.zeros(S, np.int32) for _ in range(4))
        gd = np.zeros(S)
        tp = tv = ta = None
        if record:
            tp, tv, ta = (np.zeros(3 * (max_steps + 1) * N * S) for _ in range(3))
        _lib.check(self.L.dmpcb200_run_batch(self.h, int(max_steps), int(stop_on_fail), int(mode), _p(tp), _p(tv),
                                             _p(ta), steps.ctypes.data_as(_ip), reached.ctypes.data_as(_ip),
                                             fs.ctypes.data_as(_ip), fa.ctypes.data_as(_ip), _p(gd)), "run_batch")
        ms, ast = C.c_double(0), C.c_int64(0)
        _lib.check(self.L.dmpcb200_last_batch_timing(self.h, C.byref(ms), C.byref(ast)), "last_batch_timing")
        res = dict(steps=steps, reached=reached.astype(bool), first_fail_step=fs, first_fail_agent=fa, goal_dist=gd,
                   device_ms=float(ms.value), agent_steps=int(ast.value))
        if record:
            shp = (3, max_steps + 1, N)
            blk = 3 * (max_steps + 1) * N
            for k, b in (("pk", tp), ("vk", tv), ("ak", ta)):
                res[k] = [np.asfortranarray(b[s * blk:(s + 1) * blk].reshape(shp, order="F"))[:, :steps[s] + 1, :]
                          for s in range(S)]
        return res

    def get_scenario(self, s):
        N, K = self.N, self.K
        l = np.zeros((3, K, N), order="F")
        pk, vk, ak = (np.zeros((3, N), order="F") for _ in range(3))
        status = np.zeros(N, np.int32)
        diag = np.zeros(N, dtype=_DIAG_DT)
        _lib.check(self.L.dmpcb200_get_scenario(self.h, int(s), _p(l), _p(pk), _p(vk), _p(ak),
                                                status.ctypes.data_as(_ip), diag.ctypes.data_as(C.POINTER(Diag))),
                   "get_scenario")
        return dict(l=l, pk=pk, vk=vk, ak=ak, status=status, diag=diag)

    # -- one Jacobi step, host buffers ---------------------------------------------------------------
    def step(self, pk, vk, ak, l_prev, want_horizons=False, out=None):
        """Body of `for n = 1:N` (failure_rate.m:100-119) for agents [n0,n1), HOST arrays in/out."""
        N, K = self.N, self.K
        pk, vk, ak = _f(pk, (3, N)), _f(vk, (3, N)), _f(ak, (3, N))
        l_prev = _f(l_prev, (3, K, N))
        if out is None:
            out = dict(l_new=l_prev.copy(order="F"), p1=pk.copy(order="F"), v1=vk.copy(order="F"),
                       a1=ak.copy(order="F"), status=np.zeros(N, np.int32), diag=np.zeros(N, dtype=_DIAG_DT))
            if want_horizons:
                out["v_hor"] = np.zeros((3, K, N), order="F")
                out["a_hor"] = np.zeros((3, K, N), order="F")
        ff = C.c_int32(-1)
        _lib.check(self.L.dmpcb200_step(
            self.h, _p(pk), _p(vk), _p(ak), _p(l_prev), _p(out["l_new"]), _p(out["p1"]), _p(out["v1"]),
            _p(out["a1"]), _p(out.get("v_hor")), _p(out.get("a_hor")), out["status"].ctypes.data_as(_ip),
            out["diag"].ctypes.data_as(C.POINTER(Diag)), C.byref(ff)), "step")
        out["first_fail"] = int(ff.value)
        return out

    def bind_step(self, pk, vk, ak, l_prev, out):
        """Pre-marshal one set of caller-owned HOST arrays (float64, column-major; `out` as returned by
        `step`) and return a zero-argument callable that runs `dmpcb200_step` on them -- what a tight
        closed loop over preallocated (pinned) buffers uses instead of re-deriving the pointers every step.
        Returns first_fail."""
        N, K = self.N, self.K
        for a, shp in ((pk, (3, N)), (vk, (3, N)), (ak, (3, N)), (l_prev, (3, K, N)), (out["l_new"], (3, K, N)),
                       (out["p1"], (3, N)), (out["v1"], (3, N)), (out["a1"], (3, N))):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.f_contiguous and a.shape == shp):
                raise DmpcError("bind_step: arrays must be float64, column-major, of the documented shapes")
        slot, ff = C.c_int32(-1), C.c_int32(-1)
        _lib.check(self.L.dmpcb200_bind_step(
            self.h, _p(pk), _p(vk), _p(ak), _p(l_prev), _p(out["l_new"]), _p(out["p1"]), _p(out["v1"]),
            _p(out["a1"]), _p(out.get("v_hor")), _p(out.get("a_hor")), out["status"].ctypes.data_as(_ip),
            out["diag"].ctypes.data_as(C.POINTER(Diag)), C.byref(slot)), "bind_step")
        fn, chk, h, sl, pff = self.L.dmpcb200_step_bound, _lib.check, self.h, slot.value, C.byref(ff)
        keep = (pk, vk, ak, l_prev, out)  # the arrays stay alive as long as the binding does

        def call(_keep=keep):
            rc = fn(h, sl, pff)
            if rc:
                chk(rc, "step")
            return ff.value
        return call

    def step_dev(self, d_pk, d_vk, d_ak, d_l_prev, d_l_new, d_p1, d_v1, d_a1, d_status, stream=0,
                 d_v_hor=0, d_a_hor=0, d_diag=0):
        """Same step on raw DEVICE pointers (ints), asynchronous on `stream` (cudaStream_t as int)."""
        vp = lambda x: C.c_void_p(int(x) if x else None)
        _lib.check(self.L.dmpcb200_step_dev(self.h, vp(d_pk), vp(d_vk), vp(d_ak), vp(d_l_prev), vp(d_l_new),
                                            vp(d_p1), vp(d_v1), vp(d_a1), vp(d_v_hor), vp(d_a_hor), vp(d_status),
                                            vp(d_diag), vp(stream)), "step_dev")

    def goal_dev(self, d_p, ld, d_out, stream=0):
        vp = lambda x: C.c_void_p(int(x) if x else None)
        _lib.check(self.L.dmpcb200_goal_dev(self.h, vp(d_p), int(ld), vp(d_out), vp(stream)), "goal_dev")

    def reached_goal(self, p, pf, tol):
        """ReachedGoal.m on host arrays (3,N): returns (pass, max_dist)."""
        p, pf = _f(p, (3, self.N)), _f(pf, (3, self.N))
        md, ok = C.c_double(0), C.c_int32(0)
        _lib.check(self.L.dmpcb200_reached_goal(self.h, _p(p), _p(pf), float(tol), C.byref(md), C.byref(ok)),
                   "reached_goal")
        return bool(ok.value), float(md.value)

    # -- device-resident closed loop -----------------------------------------------------------------
    def run(self, max_steps, stop_on_fail=False, mode=0, record=False, status_hist=False):
        """`while ~reached_goal && k < max_K` (failure_rate.m:99-127) on the device."""
        N = self.N
        tp = tv = ta = hist = None
        if record:
            tp, tv, ta = (np.zeros((3, max_steps + 1, N), order="F") for _ in range(3))
        if status_hist:
            hist = np.zeros((max_steps, N), np.int32)
            if not record:
                tp, tv, ta = (np.zeros((3, max_steps + 1, N), order="F") for _ in range(3))
        steps, reached, fs, fa = (C.c_int32(0) for _ in range(4))
        _lib.check(self.L.dmpcb200_run(self.h, int(max_steps), int(stop_on_fail), int(mode), _p(tp), _p(tv),
                                       _p(ta), None if hist is None else hist.ctypes.data_as(_ip),
                                       C.byref(steps), C.byref(reached), C.byref(fs), C.byref(fa)), "run")
        s = int(steps.value)
        res = dict(steps=s, reached=bool(reached.value), first_fail_step=int(fs.value),
                   first_fail_agent=int(fa.value))
        if tp is not None:
            res.update(pk=tp[:, :s + 1, :], vk=tv[:, :s + 1, :], ak=ta[:, :s + 1, :])
        if hist is not None:
            res["status_hist"] = hist[:s]
        return res

    def last_host_timing(self):
        """host-side phases of the last `step` in microseconds"""
        us = (C.c_double * 4)()
        _lib.check(self.L.dmpcb200_last_host_timing(self.h, us), "last_host_timing")
        return dict(pack_us=us[0], submit_us=us[1], wait_us=us[2], unpack_us=us[3])

    def postprocess(self, pk, vk, ak, vmax=2.0, amax=1.0, Ts=0.01, goal_radius=0.05, want_interp=True, scenario=None):
        """The rest of the reference's t_dmpc (test/failure_rate.m:134-195) for a finished transition: time
        scaling to the limits, 100 Hz spline interpolation, pairwise collision check, distance and trajectory
        time.  pk, vk, ak: (3, S, N) as `run(record=True)` returns them.  Returns a dict with the scaled
        pk, vk, ak, the figures, and (want_interp) p, v, a of shape (3, nt, N)."""
        N = self.N
        pk, vk, ak = (np.array(x, dtype=np.float64, order="F") for x in (pk, vk, ak))
        S = pk.shape[1]
        res = Post()
        tidx = np.zeros(N, np.int32)
        p = v = a = None
        cap = 0
        if want_interp:
            # size of the 100 Hz grid: T = (S-1) h / sqrt(r_factor); r_factor comes from the same norms
            with np.errstate(divide="ignore"):
                rf = min((amax / np.sqrt((ak[0] ** 2 + ak[1] ** 2) + ak[2] ** 2)).min(),
                         (vmax / np.sqrt((vk[0] ** 2 + vk[1] ** 2) + vk[2] ** 2)).min())
            cap = int(np.floor((S - 1) * self.P.h / np.sqrt(rf) / Ts + 1e-9)) + 2
            p, v, a = (np.zeros((3, cap, N), order="F") for _ in range(3))
        if scenario is None:
            _lib.check(self.L.dmpcb200_postprocess(self.h, S, _p(pk), _p(vk), _p(ak), float(vmax), float(amax), float(Ts),
                                                   float(goal_radius), _p(p), _p(v), _p(a), cap,
                                                   tidx.ctypes.data_as(_ip), C.byref(res)), "postprocess")
        else:  # goals of one scenario of a batched handle
            _lib.check(self.L.dmpcb200_postprocess_scenario(self.h, int(scenario), S, _p(pk), _p(vk), _p(ak), float(vmax),
                                                            float(amax), float(Ts), float(goal_radius), _p(p), _p(v),
                                                            _p(a), cap, tidx.ctypes.data_as(_ip), C.byref(res)),
                       "postprocess_scenario")
        out = dict(pk=pk, vk=vk, ak=ak, time_index=tidx, **{k: getattr(res, k) for k, _ in Post._fields_})
        if want_interp:
            nt = res.nt
            # the library wrote (3, nt, N) contiguously at the start of the (3, cap, N) buffers
            out.update({k: np.asfortranarray(b.ravel(order="F")[: 3 * nt * N].reshape((3, nt, N), order="F"))
                        for k, b in (("p", p), ("v", v), ("a", a))})
        return out

    def get_state(self):
        N, K = self.N, self.K
        l = np.zeros((3, K, N), order="F")
        pk, vk, ak = (np.zeros((3, N), order="F") for _ in range(3))
        status = np.zeros(N, np.int32)
        diag = np.zeros(N, dtype=_DIAG_DT)
        _lib.check(self.L.dmpcb200_get_state(self.h, _p(l), _p(pk), _p(vk), _p(ak), status.ctypes.data_as(_ip),
                                             diag.ctypes.data_as(C.POINTER(Diag))), "get_state")
        return dict(l=l, pk=pk, vk=vk, ak=ak, status=status, diag=diag)

    def set_state(self, l, pk, vk, ak):
        N, K = self.N, self.K
        l, pk, vk, ak = _f(l, (3, K, N)), _f(pk, (3, N)), _f(vk, (3, N)), _f(ak, (3, N))
        _lib.check(self.L.dmpcb200_set_state(self.h, _p(l), _p(pk), _p(vk), _p(ak)), "set_state")

    # -- per-agent drop-ins ------------------------------------------------------------------------
    def solve_agent(self, po, pf, vo, ao, n, l):
        """n 0-based.  Returns p, v, a (3,K), status, diag."""
        K = self.K
        po, pf, vo, ao = (_f(x).ravel() for x in (po, pf, vo, ao))
        l = _f(l, (3, K, self.N))
        p, v, a = (np.zeros((3, K), order="F") for _ in range(3))
        st = C.c_int32(0)
        dg = Diag()
        _lib.check(self.L.dmpcb200_solve_agent(self.h, _p(po), _p(pf), _p(vo), _p(ao), int(n), _p(l), _p(p), _p(v),
                                               _p(a), C.byref(st), C.byref(dg)), "solve_agent")
        return p, v, a, int(st.value), dict(kstar=dg.kstar, nv=dg.nv, iters=dg.iters, nact=dg.nact)

    def check_coll(self, p3, l, n, k):
        """CheckCollSoftDMPC.m; n 0-based, k 1-based.  Returns violation, min_dist, viol_constr, any."""
        N = self.N
        p3 = _f(p3).ravel()
        l = _f(l, (3, self.K, N))
        viol, vc = np.zeros(N, np.uint8), np.zeros(N, np.uint8)
        md, anyv = C.c_double(0), C.c_int32(0)
        _lib.check(self.L.dmpcb200_check_coll(self.h, _p(p3), _p(l), int(n), int(k), viol.ctypes.data_as(_u8p),
                                              vc.ctypes.data_as(_u8p), C.byref(md), C.byref(anyv)), "check_coll")
        return viol, float(md.value), vc, bool(anyv.value)

    def coll_constr(self, p3, po, vo, n, k, l, mask=None, cap=None):
        """CollConstr*DMPC.m dense rows; n 0-based, k 1-based.  Returns Ain (nv,3K), bin, prev_dist."""
        N, K = self.N, self.K
        p3, po, vo = (_f(x).ravel() for x in (p3, po, vo))
        l = _f(l, (3, K, N))
        cap = int(cap or max(N - 1, 1))
        Ain = np.zeros((cap, 3 * K), order="F")
        bin_, pd = np.zeros(cap), np.zeros(cap)
        nr = C.c_int32(0)
        m = None if mask is None else np.ascontiguousarray(np.asarray(mask).astype(np.uint8).ravel())
        _lib.check(self.L.dmpcb200_coll_constr(self.h, _p(p3), _p(po), _p(vo), int(n), int(k), _p(l),
                                               None if m is None else m.ctypes.data_as(_u8p), cap, _p(Ain), _p(bin_),
                                               _p(pd), C.byref(nr)), "coll_constr")
        r = int(nr.value)
        return np.ascontiguousarray(Ain[:r]), bin_[:r].copy(), pd[:r].copy()

    def prop_state(self, po, vo, a):
        """propStatedmpc.m for a batch: po, vo (3,B), a (3K,B) -> p, v (3K,B)."""
        K = self.K
        a = _f(a)
        a = a.reshape(3 * K, -1, order="F")
        B = a.shape[1]
        po, vo = _f(po, (3, B)), _f(vo, (3, B))
        p, v = np.zeros((3 * K, B), order="F"), np.zeros((3 * K, B), order="F")
        _lib.check(self.L.dmpcb200_prop_state(self.h, B, _p(po), _p(vo), _p(np.asfortranarray(a)), _p(p), _p(v)),
                   "prop_state")
        return p, v

    # -- introspection ------------------------------------------------------------------------------
    def last_timing(self):
        ms = (C.c_double * 3)()
        n = C.c_int64(0)
        _lib.check(self.L.dmpcb200_last_timing(self.h, ms, C.byref(n)), "last_timing")
        return dict(scan_ms=ms[0], qp_ms=ms[1], step_ms=ms[2], launches=int(n.value))

    def config(self):
        o = (C.c_int32 * 8)()
        _lib.check(self.L.dmpcb200_config(self.h, o), "config")
        keys = ("agents_per_qp_block", "QMAX", "RCAP", "RMAX", "QBIG", "rescue_slots", "qp_smem", "scan_smem")
        return dict(zip(keys, [int(x) for x in o]))

    def device_ptr(self, which: int) -> int:
        return int(self.L.dmpcb200_device_ptr(self.h, int(which)) or 0)

    def swap_horizons(self):
        _lib.check(self.L.dmpcb200_swap_horizons(self.h), "swap_horizons")


_DIAG_DT = np.dtype([("kstar", np.int32), ("nv", np.int32), ("iters", np.int32), ("nact", np.int32)])


# =================================================================================================
# Reference-named functions (dmpc/matlab).  Indices n, k are 1-based like MATLAB.
# =================================================================================================
_SOLVERS: dict = {}


def _solver_for(N, K, h, rmin, c, alim, Q1, S1, term, variant, pmin=None, pmax=None) -> Solver:
    key = (int(N), int(K), float(h), float(rmin), float(c), float(alim), float(Q1), float(S1), float(term),
           int(variant))
    s = _SOLVERS.get(key)
    if s is None:
        if len(_SOLVERS) > 8:
            _SOLVERS.pop(next(iter(_SOLVERS))).close()
        P = default_params(variant, K=int(K), h=float(h), rmin=float(rmin), c=float(c), alim=float(alim),
                           Q1=float(Q1), S1=float(S1), term=float(term))
        s = Solver(N, P)
        _SOLVERS[key] = s
    if pmin is not None:
        s.set_bounds(pmin, pmax)
    return s


def _c_from_E1(E1):
    E1 = np.asarray(E1, float)
    if E1.shape != (3, 3) or abs(E1[0, 0] - 1) > 1e-12 or abs(E1[1, 1] - 1) > 1e-12:
        raise DmpcError("E1 must be diag(1,1,1/c)")
    return 1.0 / E1[2, 2]


def _order2(order):
    if int(order) != 2:
        raise DmpcError("only order = 2 (ellipsoid) is implemented; the reference never uses another value")


def getPosMat(h, K):
    """getPosMat.m:1-23 -> A (3K x 3K)."""
    A = np.zeros((3 * K, 3 * K), order="F")
    _lib.check(_lib.lib().dmpcb200_model_mats(float(h), int(K), _p(A), None, None, None), "model_mats")
    return A


def getDeltaMat(K):
    """getDeltaMat.m:1-9 -> Delta (3K x 3K)."""
    D = np.zeros((3 * K, 3 * K), order="F")
    _lib.check(_lib.lib().dmpcb200_model_mats(1.0, int(K), None, None, None, _p(D)), "model_mats")
    return D


def modelMats(h, K):
    """A_p, A_v, A_initp, Delta of dmpc_soft_bound.m:81-108."""
    n = 3 * K
    A, Av, A0, D = (np.zeros((n, n), order="F"), np.zeros((n, n), order="F"), np.zeros((n, 6), order="F"),
                    np.zeros((n, n), order="F"))
    _lib.check(_lib.lib().dmpcb200_model_mats(float(h), int(K), _p(A), _p(Av), _p(A0), _p(D)), "model_mats")
    return A, Av, A0, D


def initDMPC(po, pf, h, k_hor, K=None):
    """initDMPC.m:1-13 -> p, v, a (3 x k_hor)."""
    po, pf = _f(po).reshape(3, 1), _f(pf).reshape(3, 1)
    with Solver(1, default_params(SOFT_BOUND, K=int(k_hor), h=float(h)), pf=pf) as s:
        l, _, _, _ = s.init_horizons(po)
    p = l[:, :, 0].copy()
    return p, np.zeros_like(p), np.zeros_like(p)


def propStatedmpc(po, vo, a, A_initp=None, A_p=None, A_v=None, h=None):
    """propStatedmpc.m:1-8 -> p, v (3K x 1).  The matrices are accepted for signature parity; the
    device evaluates the same linear maps from h (inferred from A_initp when not given)."""
    a = _f(a).ravel()
    K = a.size // 3
    if h is None:
        if A_initp is None:
            raise DmpcError("propStatedmpc needs A_initp or h")
        h = float(np.asarray(A_initp)[0, 3])
    s = _solver_for(1, K, h, 0.35, 2.0, 1.0, 1000.0, 100.0, -5e4, SOFT_BOUND)
    p, v = s.prop_state(_f(po).reshape(3, 1), _f(vo).reshape(3, 1), a.reshape(-1, 1))
    return p[:, 0], v[:, 0]


def CheckCollSoftDMPC(p, l, n, k, E1, rmin, order=2):
    """CheckCollSoftDMPC.m:1-17 -> violation (N,), min_dist, viol_constr (N,)."""
    _order2(order)
    l = _f(l)
    K, N = l.shape[1], l.shape[2]
    s = _solver_for(N, K, 0.2, rmin, _c_from_E1(E1), 1.0, 1000.0, 100.0, -5e4, SOFT_BOUND)
    viol, md, vc, _ = s.check_coll(p, l, int(n) - 1, int(k))
    return viol.astype(bool), md, vc.astype(bool)


def _coll_constr(variant, p, po, vo, n, k, l, rmin, A_initp, E1, order, violation):
    _order2(order)
    l = _f(l)
    K, N = l.shape[1], l.shape[2]
    h = float(np.asarray(A_initp)[0, 3])
    s = _solver_for(N, K, h, rmin, _c_from_E1(E1), 1.0, 1000.0, 100.0, -5e4, variant)
    return s.coll_constr(p, po, vo, int(n) - 1, int(k), l, mask=violation)


def CollConstrSoftDMPC(p, po, vo, n, k, l, rmin, Ain, A_initp, E1, E2, order, violation):
    """CollConstrSoftDMPC.m:1-32 -> Ainr, binr, prev_dist (k_ctr = k)."""
    return _coll_constr(SOFT_BOUND, p, po, vo, n, k, l, rmin, A_initp, E1, order, violation)


def CollConstrSoftDMPC2(p, po, vo, n, k, l, rmin, Ain, A_initp, E1, E2, order, violation):
    """CollConstrSoftDMPC2.m (k_ctr = k-1)."""
    return _coll_constr(SOFT_BOUND2, p, po, vo, n, k, l, rmin, A_initp, E1, order, violation)


def CollConstrHardDMPCOnDemand(p, po, vo, n, k, l, rmin, Ain, A_initp, E1, E2, order, violation):
    """CollConstrHardDMPCOnDemand.m:1-32 -> Ain_total, bin_total (two outputs, like the reference)."""
    A, b, _ = _coll_constr(HARD_ONDEMAND, p, po, vo, n, k, l, rmin, A_initp, E1, order, violation)
    return A, b


def CollConstrHardDMPC(p, po, vo, n, k, l, rmin, Ain, A_initp, E1, E2, order):
    """CollConstrHardDMPC.m:1-34 -> Ain_total (N-1, 3K), bin_total (N-1,): one row per neighbour with
    dist < 1 (:19), packed first in ascending neighbour order, and all-zero rows for the rest (:3-5) --
    the reference's exact shape (solveHardDMPC.m:18-22 stacks K of these blocks).  The device returns the
    non-trivial rows; the zero rows are padding added here."""
    A, b, _ = _coll_constr(HARD, p, po, vo, n, k, l, rmin, A_initp, E1, order, None)
    N, K = np.asarray(l).shape[2], np.asarray(l).shape[1]
    Ain_total, bin_total = np.zeros((max(N - 1, 0), 3 * K)), np.zeros(max(N - 1, 0))
    Ain_total[:A.shape[0]] = A
    bin_total[:A.shape[0]] = b
    return Ain_total, bin_total


def CollConstr(p, po, k, l, Ain, rmin, E1, E2, order):
    """The dec-iSCP helper name (dec-iSCP/CollConstr.m:1-23) on the device path -> Ain_total (N_obs, 3K),
    bin_total (N_obs,).  Its rows are the DMPC rows of CollConstrSoftDMPC2 with vo = 0 and no own agent: every
    column of l is an obstacle, the row acts on block k-1 (:17), r = dist (rmin - dist + diff p / dist) -
    diff po' (:14).  Ain must be getPosMat(h, K) (dec-iSCP/singleiSCP.m:9); h is read from it."""
    _order2(order)
    if int(k) < 2:
        raise DmpcError("CollConstr needs k >= 2 (the row acts on block k-1)")
    l = _f(l)
    K, N = l.shape[1], l.shape[2]
    h = float(np.sqrt(2.0 * np.asarray(Ain)[0, 0]))
    s = _solver_for(N, K, h, rmin, _c_from_E1(E1), 1.0, 1000.0, 100.0, -5e4, SOFT_BOUND2)
    A, b, _ = s.coll_constr(p, po, np.zeros(3), -1, int(k), l, mask=np.ones(N, np.uint8), cap=N)
    return A, b


def propState(po, a, A_p, A_v, K):
    """The dec-iSCP helper name (dec-iSCP/propState.m:1-10) on the device path: K trajectory points from rest,
    p = [po; A_p a + po], v = [0; A_v a] with A_p, A_v the first 3(K-1) rows of the kinematic maps
    (dec-iSCP/decSCP.m:55-71), i.e. the first K-1 rows of propStatedmpc.m with vo = 0."""
    K = int(K)
    h = float(np.asarray(A_v)[0, 0])
    pp, vv = propStatedmpc(po, np.zeros(3), a, h=h)
    po = _f(po).ravel()
    return np.r_[po, pp[:3 * (K - 1)]], np.r_[np.zeros(3), vv[:3 * (K - 1)]]


def is_inbounds(p, pmin, pmax, tol=50e-3):
    """is_inbounds.m:1-6 (host scalar test, as in the reference)."""
    p = np.asarray(p, float).reshape(3, -1)
    pmin, pmax = np.asarray(pmin, float).ravel(), np.asarray(pmax, float).ravel()
    return bool(np.all(p.max(1) < pmax + tol) and np.all(p.min(1) > pmin - tol))


def ReachedGoal(p, pf, length_t, error_tol, N):
    """ReachedGoal.m:1-11: p (3, T, N), pf (3, N) or (1,3,N); length_t 1-based."""
    p = np.asarray(p, float)
    pk = p[:, int(length_t) - 1, :] if p.ndim == 3 else p[:, int(length_t) - 1].reshape(3, 1)
    pf = np.asarray(pf, float).reshape(3, -1, order="F")
    s = _solver_for(int(N), 15, 0.2, 0.35, 2.0, 1.0, 1000.0, 100.0, -5e4, SOFT_BOUND)
    return s.reached_goal(pk, pf, error_tol)[0]


def status_flags(st: int, variant: int):
    """status word -> (solved, success/feasible, outbound, coll) exactly as the reference function of the
    variant returns them.  solveSoftDMPCbound.m keeps `feasible = 1` on the k == 1 collision exit (:29-30)
    and when the first position is out of bounds (:125-128 only sets `outbound`); solveSoftDMPCbound2.m
    (:14,26,123-126), solveHardDMPC.m (:14,76-79) and solveHardDMPCOnDemand.m (:14,81-84) start from
    `success = 0`, leave it 0 on the collision exit and reset it to 0 when out of bounds.  The library's
    internal failures (iteration cap, capacity overflow) count as not feasible for every variant."""
    st = int(st)
    solved = bool(st & ST_SOLVED)
    outbound = 1 if (st & ST_OUTBOUND) else 0
    coll = 1 if (st & ST_COLL) else 0
    hard_fail = bool(st & (ST_INFEASIBLE | ST_QPFAIL | ST_OVERFLOW))
    if int(variant) == SOFT_BOUND:
        success = 0 if hard_fail else 1
    else:
        success = 1 if (solved and not outbound and not coll and not hard_fail) else 0
    return solved, success, outbound, coll


def _solve(variant, po, pf, vo, ao, n, h, l, K, rmin, pmin, pmax, alim, Q1, S1, E1, order, term):
    _order2(order)
    l = _f(l)
    N = l.shape[2]
    s = _solver_for(N, K, h, rmin, _c_from_E1(E1), alim, Q1, S1, term, variant, pmin, pmax)
    p, v, a, st, dg = s.solve_agent(po, pf, vo, ao, int(n) - 1, l)
    solved, feasible, outbound, coll = status_flags(st, variant)
    if not solved:
        e = np.zeros((0, 0))
        return e, e, e, feasible, outbound, coll
    return p, v, a, feasible, outbound, coll


def solveSoftDMPCbound(po, pf, vo, ao, n, h, l, K, rmin, pmin, pmax, alim, A, A_initp, A_p, A_v, Delta, Q1, S1, E1,
                       E2, order, term):
    """solveSoftDMPCbound.m:1-160 -> [p, v, a, feasible, outbound, coll]; p, v, a are (3,K) or empty."""
    return _solve(SOFT_BOUND, po, pf, vo, ao, n, h, l, K, rmin, pmin, pmax, alim, Q1, S1, E1, order, term)


def solveSoftDMPCbound2(po, pf, vo, ao, n, h, l, K, rmin, pmin, pmax, alim, A, A_initp, A_p, A_v, Delta, Q1, S1,
                        E1, E2, order, term):
    """solveSoftDMPCbound2.m:1-149."""
    return _solve(SOFT_BOUND2, po, pf, vo, ao, n, h, l, K, rmin, pmin, pmax, alim, Q1, S1, E1, order, term)


def solveHardDMPC(po, pf, vo, ao, n, h, l, K, rmin, pmin, pmax, alim, A, A_initp, A_p, A_v, Delta, Q1, S1, E1, E2,
                  order):
    """solveHardDMPC.m:1-90 -> [p, v, a, success, outbound, coll]."""
    p, v, a, feas, outb, coll = _solve(HARD, po, pf, vo, ao, n, h, l, K, rmin, pmin, pmax, alim, Q1, S1, E1, order,
                                       -5e4)
    return p, v, a, feas, outb, coll


def solveHardDMPCOnDemand(po, pf, vo, ao, n, h, l, K, rmin, pmin, pmax, alim, A, A_initp, A_p, A_v, Delta, Q1, S1,
                          E1, E2, order):
    """solveHardDMPCOnDemand.m:1-95."""
    return _solve(HARD_ONDEMAND, po, pf, vo, ao, n, h, l, K, rmin, pmin, pmax, alim, Q1, S1, E1, order, -5e4)


def close_cached_solvers():
    for s in list(_SOLVERS.values()):
        s.close()
    _SOLVERS.clear()


# =================================================================================================
# C++ mirror: class DMPC (dmpc/cpp/dmpc.h:70-182)
# =================================================================================================
class DMPC:
    """Mirror of the C++ `class DMPC` public surface for the parallel solve
    (set_boundaries / set_initial_pts / set_final_pts / set_k_factor / set_cluster_num /
    solveParallelDMPCv2, dmpc.h:83-172).  `set_cluster_num` is accepted and ignored: the agents are
    the warps of one B200 (or the ranks' shards), not host threads."""

    def __init__(self, params: dict | None = None):
        d = dict(h=0.2, T=30.0, k_hor=15, order=2, c=2.0, rmin=0.35, alim=1.0, vlim=2.0, freq=100.0,
                 goal_tol=0.01, collision_tol=0.05, speed=1)
        d.update(params or {})
        _order2(d["order"])
        self.params = d
        self.k_factor = 0
        self.pmin = np.array([-2.5, -2.5, 0.2])
        self.pmax = np.array([2.5, 2.5, 2.2])
        self.po = self.pf = None
        self.successful = False
        self.solution_short = None

    def set_boundaries(self, pmin, pmax):
        self.pmin, self.pmax = np.asarray(pmin, float).ravel(), np.asarray(pmax, float).ravel()

    def set_initial_pts(self, po):
        self.po = _f(po).reshape(3, -1, order="F")

    def set_final_pts(self, pf):
        self.pf = _f(pf).reshape(3, -1, order="F")

    def set_k_factor(self, k_factor):
        if k_factor not in (0, -1):
            raise DmpcError("k_factor must be 0 or -1 (dmpc.cpp:516)")
        self.k_factor = int(k_factor)

    def set_cluster_num(self, n):
        self.n_clusters = int(n)

    def solveParallelDMPCv2(self, stop_on_fail=True):
        """dmpc.cpp:1570-1740 up to the end of the MPC loop (post-processing is out of scope).
        Returns a list of per-agent dicts {pos, vel, acc} (3 x steps) like std::vector<Trajectory>."""
        d = self.params
        N = self.po.shape[1]
        variant = SOFT_BOUND if self.k_factor == 0 else SOFT_BOUND2
        P = default_params(variant, K=int(d["k_hor"]), h=float(d["h"]), rmin=float(d["rmin"]), c=float(d["c"]),
                           alim=float(d["alim"]), goal_tol=float(d["goal_tol"]), coll_tol=float(d["collision_tol"]))
        max_K = int(d["T"] / d["h"])
        with Solver(N, P, pmin=self.pmin, pmax=self.pmax, pf=self.pf) as s:
            s.init_horizons(self.po)
            r = s.run(max_K - 1, stop_on_fail=stop_on_fail, record=True)
        self.successful = bool(r["reached"]) and r["first_fail_step"] < 0
        self.solution_short = [dict(pos=r["pk"][:, :, n], vel=r["vk"][:, :, n], acc=r["ak"][:, :, n])
                               for n in range(N)]
        self.steps = r["steps"]
        return self.solution_short
