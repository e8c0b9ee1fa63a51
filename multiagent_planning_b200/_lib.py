"""Loader / builder of libdmpc_b200.so (the C-ABI of include/dmpc_b200.h) and its ctypes prototypes.

The library is hand-written CUDA for sm_100a (multiagent_planning_b200/csrc).  It is built IN-TREE
with nvcc (``build()``), never JIT-compiled, and there is no fallback: if it is missing or no B200
is present, compute calls raise.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libdmpc_b200.so")
_ALT = os.environ.get("DMPCB200_LIB")  # experiment hook: an alternative build of the library (A/B runs, profiling builds)
SOURCES = ["dmpc_b200.cu", "k_qp15.cu", "k_qp20.cu", "k_qpgen.cu", "k_qphard.cu", "k_scan.cu",
           "model_tables.cpp", "traj_io.cpp"]
HEADERS = ["dmpc_kernels.cuh", "small_kernels.cuh", "launch.cuh", "scan_core.cuh", "agent_solve.cuh", "qp_core.cuh",
           "qp_warp.cuh", "postprocess.cuh", "model_tables.h", os.path.join("..", "..", "include", "dmpc_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"]
HASH_PATH = LIB_PATH + ".srchash"

_LIB = None


class DmpcError(RuntimeError):
    """API / CUDA error reported by libdmpc_b200 (negative return code)."""


class Params(C.Structure):
    """dmpcb200_params (include/dmpc_b200.h) -- mirrors `struct Params` of dmpc/cpp/dmpc.h:50-63 plus
    the constants the MATLAB scripts hard-code."""
    _fields_ = [
        ("K", C.c_int32), ("variant", C.c_int32), ("max_tries", C.c_int32), ("neigh_mode", C.c_int32),
        ("h", C.c_double), ("rmin", C.c_double), ("c", C.c_double), ("alim", C.c_double),
        ("Q1", C.c_double), ("S1", C.c_double), ("term", C.c_double),
        ("Q_far", C.c_double), ("Q_near", C.c_double), ("S_free", C.c_double),
        ("near_radius", C.c_double), ("slack_lb", C.c_double), ("neigh_factor", C.c_double),
        ("coll_tol", C.c_double), ("inb_tol", C.c_double), ("hard_radius", C.c_double),
        ("init_div", C.c_double), ("goal_tol", C.c_double),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Post(C.Structure):
    """dmpcb200_post (include/dmpc_b200.h): figures of test/failure_rate.m:134-195"""
    _fields_ = [("r_factor", C.c_double), ("h_scaled", C.c_double), ("T", C.c_double), ("min_dist", C.c_double),
                ("totdist", C.c_double), ("traj_time", C.c_double), ("nt", C.c_int32), ("violation", C.c_int32),
                ("device_ms", C.c_double)]


class Diag(C.Structure):
    _fields_ = [("kstar", C.c_int32), ("nv", C.c_int32), ("iters", C.c_int32), ("nact", C.c_int32)]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise DmpcError("nvcc not found: cannot build libdmpc_b200.so")


def source_hash() -> str:
    """sha256 over every source, header and compiler flag of the library.  The built .so carries it in a
    side file: `build()` rebuilds whenever the tree and the binary disagree (mtimes are not trusted: a
    binary that travelled to another box, or a checkout, has arbitrary times)."""
    hh = hashlib.sha256()
    hh.update(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        path = os.path.join(_CSRC, f)
        hh.update(f.encode())
        with open(path, "rb") as fh:
            hh.update(fh.read())
    return hh.hexdigest()


def built_hash() -> str | None:
    try:
        with open(HASH_PATH) as fh:
            return fh.read().strip()
    except OSError:
        return None


def is_stale() -> bool:
    return not os.path.exists(LIB_PATH) or built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> multiagent_planning_b200/libdmpc_b200.so
    (one object per translation unit, compiled in parallel, then linked)."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    want = source_hash()
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        procs.append((src, obj, subprocess.Popen(cmd, cwd=_CSRC)))
    objs = []
    for src, obj, pr in procs:
        if pr.wait() != 0:
            raise DmpcError(f"nvcc failed on {src}")
        objs.append(obj)
    if os.path.exists(HASH_PATH):
        os.remove(HASH_PATH)
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs,
                          cwd=_CSRC)
    with open(HASH_PATH, "w") as fh:
        fh.write(want + "\n")
    return LIB_PATH


def lib():
    """The loaded library with prototypes set.  Raises DmpcError when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if _ALT:
        path = _ALT
    else:
        path = LIB_PATH
    if not os.path.exists(path):
        raise DmpcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(path)
    dp, ip, u8p, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_void_p
    PP, DP = C.POINTER(Params), C.POINTER(Diag)
    I, D = C.c_int, C.c_double
    protos = {
        "dmpcb200_abi_version": ([], I),
        "dmpcb200_last_error": ([], C.c_char_p),
        "dmpcb200_device_count": ([], I),
        "dmpcb200_default_params": ([PP, I], None),
        "dmpcb200_default_params_cpp": ([PP, I], None),
        "dmpcb200_set_static_obstacles": ([vp, I], I),
        "dmpcb200_model_mats": ([D, I, dp, dp, dp, dp], I),
        "dmpcb200_create": ([PP, I, I, I, I, I, I, C.POINTER(vp)], I),
        "dmpcb200_set_scenario": ([vp, I, dp, dp, dp, dp], I),
        "dmpcb200_gen_scenarios": ([vp, C.c_uint64, I, D, dp, dp, dp, dp], I),
        "dmpcb200_run_batch": ([vp, I, I, I, dp, dp, dp, ip, ip, ip, ip, dp], I),
        "dmpcb200_get_scenario": ([vp, I, dp, dp, dp, dp, ip, DP], I),
        "dmpcb200_last_batch_timing": ([vp, dp, C.POINTER(C.c_int64)], I),
        "dmpcb200_destroy": ([vp], None),
        "dmpcb200_set_bounds": ([vp, dp, dp], I),
        "dmpcb200_set_goals": ([vp, dp], I),
        "dmpcb200_init_horizons": ([vp, dp, dp, dp, dp, dp], I),
        "dmpcb200_step": ([vp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, ip, DP, ip], I),
        "dmpcb200_bind_step": ([vp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, ip, DP, ip], I),
        "dmpcb200_step_bound": ([vp, C.c_int32, ip], I),
        "dmpcb200_step_dev": ([vp] + [vp] * 12 + [vp], I),
        "dmpcb200_goal_dev": ([vp, vp, I, vp, vp], I),
        "dmpcb200_reached_goal": ([vp, dp, dp, D, dp, ip], I),
        "dmpcb200_run": ([vp, I, I, I, dp, dp, dp, ip, ip, ip, ip, ip], I),
        "dmpcb200_get_state": ([vp, dp, dp, dp, dp, ip, DP], I),
        "dmpcb200_set_state": ([vp, dp, dp, dp, dp], I),
        "dmpcb200_solve_agent": ([vp, dp, dp, dp, dp, I, dp, dp, dp, dp, ip, DP], I),
        "dmpcb200_check_coll": ([vp, dp, dp, I, I, u8p, u8p, dp, ip], I),
        "dmpcb200_coll_constr": ([vp, dp, dp, dp, I, I, dp, u8p, I, dp, dp, dp, ip], I),
        "dmpcb200_prop_state": ([vp, I, dp, dp, dp, dp, dp], I),
        "dmpcb200_postprocess": ([vp, I, dp, dp, dp, D, D, D, D, dp, dp, dp, I, ip, C.POINTER(Post)], I),
        "dmpcb200_postprocess_scenario": ([vp, I, I, dp, dp, dp, D, D, D, D, dp, dp, dp, I, ip, C.POINTER(Post)], I),
        "dmpcb200_last_timing": ([vp, dp, C.POINTER(C.c_int64)], I),
        "dmpcb200_write_trajectories": ([C.c_char_p, I, I, I, D, dp, dp, dp, dp, dp, dp, dp], I),
        "dmpcb200_read_trajectories": ([C.c_char_p, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp], I),
        "dmpcb200_format_matrix": ([I, I, dp, C.c_char_p, I], I),
        "dmpcb200_last_host_timing": ([vp, dp], I),
        "dmpcb200_device_ptr": ([vp, I], vp),
        "dmpcb200_swap_horizons": ([vp], I),
        "dmpcb200_config": ([vp, ip], I),
    }
    for name, (args, res) in protos.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    L._protos = protos
    _LIB = L
    return L


EXPORTS = [
    "dmpcb200_abi_version", "dmpcb200_last_error", "dmpcb200_device_count", "dmpcb200_default_params",
    "dmpcb200_default_params_cpp", "dmpcb200_set_static_obstacles",
    "dmpcb200_model_mats", "dmpcb200_create", "dmpcb200_set_scenario", "dmpcb200_gen_scenarios", "dmpcb200_run_batch", "dmpcb200_get_scenario",
    "dmpcb200_last_batch_timing", "dmpcb200_destroy", "dmpcb200_set_bounds", "dmpcb200_set_goals",
    "dmpcb200_init_horizons", "dmpcb200_step", "dmpcb200_bind_step", "dmpcb200_step_bound", "dmpcb200_step_dev", "dmpcb200_goal_dev", "dmpcb200_reached_goal", "dmpcb200_run",
    "dmpcb200_get_state", "dmpcb200_set_state", "dmpcb200_solve_agent", "dmpcb200_check_coll",
    "dmpcb200_coll_constr", "dmpcb200_prop_state", "dmpcb200_postprocess", "dmpcb200_postprocess_scenario", "dmpcb200_last_timing", "dmpcb200_last_host_timing",
    "dmpcb200_device_ptr",
    "dmpcb200_swap_horizons", "dmpcb200_config",
    "dmpcb200_write_trajectories", "dmpcb200_read_trajectories", "dmpcb200_format_matrix",
]


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().dmpcb200_last_error()
        raise DmpcError(f"{what}: rc={rc}: {msg.decode() if msg else ''}")
