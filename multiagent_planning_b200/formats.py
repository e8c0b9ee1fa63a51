"""On-disk formats of the reference, so that its readers and plot scripts keep working on our output.

* ``trajectories2file`` / ``read_trajectories``: the text dump of ``DMPC::trajectories2file``
  (dmpc/cpp/dmpc.cpp:2088-2126) that ``dmpc/cpp_results/read_result.m`` reads with ``dlmread`` -- written and
  parsed by the library (``dmpcb200_write_trajectories`` / ``dmpcb200_read_trajectories``, host code), byte for
  byte Eigen's stream format.
* ``save_workspace`` / ``load_workspace``: a MATLAB ``.mat`` file with the variable names and shapes of the
  workspaces the reference's experiment scripts save (``save(...)`` at test/failure_rate.m:205:
  ``po``/``pf`` 1 x 3 x N, ``pk``/``vk``/``ak`` 3 x T x N, ``l`` 3 x K x N, the model matrices ``A``,
  ``A_p_dmpc``, ``A_v_dmpc``, ``A_initp``, ``Delta`` ...), so that ``plot/*.m`` can ``load`` a run of this
  library.  The same files are what ``tests/golden/make_golden.py`` reads from the reference's ``data/``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_dp = C.POINTER(C.c_double)


def _f(a, shape=None):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = np.asfortranarray(a.reshape(shape, order="F"))
    return a


def trajectories2file(path, po, pf, pos, vel, acc, h_scaled, pmin, pmax):
    """po (3,N), pf (3,N_cmd), pos/vel/acc (3,T,N_cmd) -> the reference's trajectories.txt"""
    po, pf = _f(po), _f(pf)
    N, Nc = po.shape[1], pf.shape[1]
    pos, vel, acc = _f(pos), _f(vel), _f(acc)
    T = pos.shape[1]
    if pos.shape != (3, T, Nc) or vel.shape != pos.shape or acc.shape != pos.shape:
        raise _lib.DmpcError("trajectories2file: pos / vel / acc must be (3, T, N_cmd)")
    pmin, pmax = _f(pmin).ravel(), _f(pmax).ravel()
    p = lambda a: a.ctypes.data_as(_dp)
    rc = _lib.lib().dmpcb200_write_trajectories(str(path).encode(), N, Nc, T, float(h_scaled), p(pmin), p(pmax), p(po),
                                                p(pf), p(pos), p(vel), p(acc))
    if rc:
        raise _lib.DmpcError(f"trajectories2file: cannot write {path} (rc={rc})")


def read_trajectories(path):
    """the inverse: dict(N, N_cmd, T, h_scaled, pmin, pmax, po, pf, pos, vel, acc)"""
    L = _lib.lib()
    n, nc, t = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    rc = L.dmpcb200_read_trajectories(str(path).encode(), C.byref(n), C.byref(nc), C.byref(t), None, None, None, None,
                                      None, None, None, None)
    if rc:
        raise _lib.DmpcError(f"read_trajectories: cannot read {path} (rc={rc})")
    N, Nc, T = n.value, nc.value, t.value
    hs = C.c_double(0)
    pmin, pmax = np.zeros(3), np.zeros(3)
    po, pf = np.zeros((3, N), order="F"), np.zeros((3, Nc), order="F")
    pos, vel, acc = (np.zeros((3, T, Nc), order="F") for _ in range(3))
    p = lambda a: a.ctypes.data_as(_dp)
    rc = L.dmpcb200_read_trajectories(str(path).encode(), C.byref(n), C.byref(nc), C.byref(t), C.byref(hs), p(pmin),
                                      p(pmax), p(po), p(pf), p(pos), p(vel), p(acc))
    if rc:
        raise _lib.DmpcError(f"read_trajectories: malformed file {path} (rc={rc})")
    return dict(N=N, N_cmd=Nc, T=T, h_scaled=float(hs.value), pmin=pmin, pmax=pmax, po=po, pf=pf, pos=pos, vel=vel,
                acc=acc)


def save_workspace(path, pk, vk, ak, po, pf, pmin, pmax, h=0.2, k_hor=15, l=None, **extra):
    """MATLAB workspace of a finished transition (variable names of test/failure_rate.m / dmpc_soft_bound.m)"""
    import scipy.io
    from . import dmpc
    po, pf = _f(po), _f(pf)
    N = po.shape[1]
    A, Av, A0, D = dmpc.modelMats(float(h), int(k_hor))
    ws = dict(pk=_f(pk), vk=_f(vk), ak=_f(ak), po=po.reshape(1, 3, N, order="F"), pf=pf.reshape(1, 3, N, order="F"),
              pmin=_f(pmin).reshape(1, 3), pmax=_f(pmax).reshape(1, 3), h=float(h), k_hor=float(k_hor), N=float(N),
              A=A, A_p_dmpc=A, A_v_dmpc=Av, A_initp=A0, Delta=D)
    if l is not None:
        ws["l"] = _f(l)
    ws.update(extra)
    scipy.io.savemat(str(path), ws)


def load_workspace(path):
    """the inverse (also reads the reference's own data/**/*.mat): po / pf come back as (3, N)"""
    import scipy.io
    raw = scipy.io.loadmat(str(path))
    out = {}
    for k, v in raw.items():
        if k.startswith("__"):
            continue
        v = np.asarray(v)
        if k in ("po", "pf") and v.ndim == 3:
            v = np.asfortranarray(v.reshape(3, -1, order="F"))
        elif v.size == 1:
            x = float(v.ravel()[0])
            v = int(x) if k in ("N", "k_hor", "K") else x
        out[k] = v
    return out
