"""The reference's trajectory text format (DMPC::trajectories2file, dmpc/cpp/dmpc.cpp:2088-2126; read by
dmpc/cpp_results/read_result.m): writer byte-exact against an excerpt of the reference's own 200-agent dump,
write -> read round trip, .mat workspace export / import."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fmt(L, m):
    m = np.asfortranarray(m, dtype=np.float64)
    n = L.dmpcb200_format_matrix(m.shape[0], m.shape[1], m.ctypes.data_as(C.POINTER(C.c_double)), None, 0)
    buf = C.create_string_buffer(n + 1)
    L.dmpcb200_format_matrix(m.shape[0], m.shape[1], m.ctypes.data_as(C.POINTER(C.c_double)), buf, n + 1)
    return buf.value.decode()


def test_matrix_format_is_eigens_byte_for_byte():
    from multiagent_planning_b200 import _lib
    L = _lib.lib()
    lines = open(os.path.join(ROOT, "tests", "golden", "traj200_excerpt.txt")).read().split("\n")
    assert lines[0].split()[:3] == ["200", "200", "0.2"]
    for start in (1, 4, 7, 10, 13):          # po, pf, first agent's pos / vel / acc: 3 lines each
        block = lines[start:start + 3]
        m = np.array([[float(x) for x in ln.split()] for ln in block])
        assert _fmt(L, m) == "\n".join(block), f"block at line {start + 1}"
    # the header's row vectors
    assert _fmt(L, np.array([[-2.5, -2.5, 0.2]])) == "-2.5 -2.5  0.2" and _fmt(L, np.array([[2.5, 2.5, 5.2]])) == "2.5 2.5 5.2"


def test_write_read_round_trip(tmp_path):
    from multiagent_planning_b200 import formats
    rng = np.random.default_rng(0)
    N, Nc, T = 7, 5, 11
    po, pf = rng.uniform(-2, 2, (3, N)), rng.uniform(-2, 2, (3, Nc))
    pos, vel, acc = (rng.normal(size=(3, T, Nc)) for _ in range(3))
    path = str(tmp_path / "trajectories.txt")
    formats.trajectories2file(path, po, pf, pos, vel, acc, h_scaled=0.1834, pmin=[-2.5, -2.5, 0.2], pmax=[2.5, 2.5, 2.2])
    r = formats.read_trajectories(path)
    assert (r["N"], r["N_cmd"], r["T"]) == (N, Nc, T) and abs(r["h_scaled"] - 0.1834) < 1e-12
    for k, v in (("po", po), ("pf", pf), ("pos", pos), ("vel", vel), ("acc", acc)):
        assert r[k].shape == v.shape and np.allclose(r[k], v, rtol=5e-6, atol=0)      # 6 significant digits
    first = open(path).readline().split()
    assert first[:2] == ["7", "5"] and len(first) == 9
    # second write of what was read: the text is a fixed point of write . read
    formats.trajectories2file(str(tmp_path / "again.txt"), r["po"], r["pf"], r["pos"], r["vel"], r["acc"],
                              h_scaled=r["h_scaled"], pmin=r["pmin"], pmax=r["pmax"])
    assert open(path).read() == open(str(tmp_path / "again.txt")).read()
    with pytest.raises(Exception):
        formats.read_trajectories(str(tmp_path / "missing.txt"))


@pytest.mark.skipif(not os.path.exists("/root/reference/dmpc/cpp_results"), reason="reference tree not present")
def test_reference_dump_is_reproduced_byte_for_byte(tmp_path):
    """read the reference's own 200-agent dump and write it again: identical bytes (where the reference is mounted)"""
    from multiagent_planning_b200 import formats
    src = "/root/reference/dmpc/cpp_results/trajectories (200-agents).txt"
    r = formats.read_trajectories(src)
    assert (r["N"], r["N_cmd"], r["T"]) == (200, 200, 83)
    out = str(tmp_path / "t.txt")
    formats.trajectories2file(out, r["po"], r["pf"], r["pos"], r["vel"], r["acc"], h_scaled=r["h_scaled"],
                              pmin=r["pmin"], pmax=r["pmax"])
    assert open(out).read() == open(src).read()


def test_mat_workspace_round_trip(tmp_path):
    """.mat export with the variable names of the reference's saved workspaces (test/failure_rate.m:205)"""
    from multiagent_planning_b200 import formats
    rng = np.random.default_rng(1)
    N, T, K = 6, 9, 15
    ws = dict(pk=rng.normal(size=(3, T, N)), vk=rng.normal(size=(3, T, N)), ak=rng.normal(size=(3, T, N)),
              po=rng.normal(size=(3, N)), pf=rng.normal(size=(3, N)), l=rng.normal(size=(3, K, N)),
              pmin=np.array([-2.5, -2.5, 0.2]), pmax=np.array([2.5, 2.5, 2.2]), h=0.2, k_hor=K, rmin=0.35, c=2.0)
    path = str(tmp_path / "ws.mat")
    formats.save_workspace(path, **ws)
    r = formats.load_workspace(path)
    assert r["po"].shape == (3, N) and r["pk"].shape == (3, T, N) and r["N"] == N and r["k_hor"] == K
    for k in ("pk", "vk", "ak", "po", "pf", "l"):
        assert np.array_equal(r[k], ws[k])
    import scipy.io
    raw = scipy.io.loadmat(path)
    assert raw["po"].shape == (1, 3, N) and raw["pf"].shape == (1, 3, N)       # MATLAB's 1 x 3 x N convention
    assert raw["A"].shape == (3 * K, 3 * K) and raw["Delta"].shape == (3 * K, 3 * K)
