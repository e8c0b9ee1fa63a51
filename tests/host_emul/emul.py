"""ctypes binding of the TEST-ONLY host build of the device algorithm (tests/host_emul/emul.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = None


class Params(C.Structure):
    """dmpcb200_params of include/dmpc_b200.h"""
    _fields_ = [
        ("K", C.c_int32), ("variant", C.c_int32), ("max_tries", C.c_int32), ("neigh_mode", C.c_int32),
        ("h", C.c_double), ("rmin", C.c_double), ("c", C.c_double), ("alim", C.c_double),
        ("Q1", C.c_double), ("S1", C.c_double), ("term", C.c_double),
        ("Q_far", C.c_double), ("Q_near", C.c_double), ("S_free", C.c_double),
        ("near_radius", C.c_double), ("slack_lb", C.c_double), ("neigh_factor", C.c_double),
        ("coll_tol", C.c_double), ("inb_tol", C.c_double), ("hard_radius", C.c_double),
        ("init_div", C.c_double), ("goal_tol", C.c_double),
    ]


def build(force=False):
    so = os.path.join(_HERE, "libemul.so")
    srcs = [os.path.join(_HERE, "emul.cpp"),
            os.path.join(_ROOT, "multiagent_planning_b200/csrc/model_tables.cpp")]
    deps = srcs + [os.path.join(_ROOT, "multiagent_planning_b200/csrc", f)
                   for f in ("qp_core.cuh", "qp_warp.cuh", "agent_solve.cuh", "scan_core.cuh", "model_tables.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared"] + os.environ.get("EMUL_FLAGS", "").split() + [
                               "-o", so] + srcs)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def params_from(obj) -> Params:
    """copy the common fields of any ctypes params struct (oracle's or the product's)"""
    P = Params()
    for name, _ in Params._fields_:
        if hasattr(obj, name):
            setattr(P, name, getattr(obj, name))
    if P.goal_tol == 0:
        P.goal_tol = 0.01
    return P


def step(P, pk, vk, ak, pf, l_prev, pmin, pmax, n0=0, n1=None, QMAX=64, RCAP=64, RMAX=None, warm=None):
    f = lambda a: np.asfortranarray(np.asarray(a, np.float64))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    l_prev = f(l_prev)
    K, N = l_prev.shape[1], l_prev.shape[2]
    n1 = N if n1 is None else n1
    RMAX = RMAX or max(N - 1, 1) * (K if P.variant == 2 else 1)
    pk, vk, ak, pf = f(pk), f(vk), f(ak), f(pf)
    l_new = l_prev.copy(order="F")
    p1, v1, a1 = pk.copy(order="F"), vk.copy(order="F"), ak.copy(order="F")
    v_hor = np.zeros_like(l_new)
    a_hor = np.zeros_like(l_new)
    status = np.zeros(N, np.int32)
    diag = np.zeros((N, 4), np.int32)
    pmin, pmax = f(pmin).ravel(), f(pmax).ravel()
    lib().emul_step(C.byref(P), N, n0, n1, dp(pk), dp(vk), dp(ak), dp(pf), dp(l_prev), dp(pmin), dp(pmax),
                    QMAX, RCAP, RMAX, dp(l_new), dp(p1), dp(v1), dp(a1), dp(v_hor), dp(a_hor),
                    status.ctypes.data_as(C.POINTER(C.c_int32)), diag.ctypes.data_as(C.POINTER(C.c_int32)),
                    warm.ctypes.data_as(C.POINTER(C.c_int32)) if warm is not None else None)
    return dict(l_new=l_new, p1=p1, v1=v1, a1=a1, v_hor=v_hor, a_hor=a_hor, status=status, diag=diag)
