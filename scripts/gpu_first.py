"""First GPU bring-up: KAT parity vs the oracle, closed-loop parity, timings.  Run under gpurun."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import dmpc, scenarios
from oracle import dmpc_oracle as orc

def oparams(P):
    O = orc.default_params(P.variant)
    for n, _ in O._fields_:
        setattr(O, n, getattr(P, n))
    return O

g = np.load("tests/golden/kat_soft_bound.npz")
N = int(g["N"])
P = dmpc.default_params(0)
s = dmpc.Solver(N, P, pmin=g["pmin"], pmax=g["pmax"], pf=g["pf"])
print("config", s.config())
out = s.step(g["pk_prev"], g["vk_prev"], g["ak_prev"], g["l"], want_horizons=True)
o = orc.step(oparams(P), g["pk_prev"], g["vk_prev"], g["ak_prev"], g["pf"], g["l"], g["pmin"], g["pmax"], want_diag=True)
print("KAT status equal", np.array_equal(out["status"] & 0xff, o["status"] & 0xff), np.bincount(out["status"] & 0xff).nonzero())
print("KAT max |l_new - oracle|", np.abs(out["l_new"] - o["l_new"]).max(), "first_fail", out["first_fail"], o["first_fail"])
ns = int(g["n_solved"])
print("KAT vs MATLAB max", np.abs(out["l_new"][:, :, :ns] - g["new_l"]).max())
print("timing", s.last_timing())
s.close()

for name in ("N100", "N500"):
    cfg = scenarios.config(name)
    N = cfg["N"]
    P = dmpc.default_params(cfg["variant"], **cfg["params"])
    s = dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"])
    s.init_horizons(cfg["po"])
    t = time.time()
    r = s.run(cfg["max_steps"], record=True, status_hist=True)
    print(name, "graph run steps", r["steps"], "reached", r["reached"], "fail", r["first_fail_step"], r["first_fail_agent"],
          "wall %.3f s" % (time.time() - t), s.last_timing())
    # closed-loop oracle replay, teacher forced per step on the GPU trajectory
    O = oparams(P)
    l, pk, vk, ak = s.init_horizons(cfg["po"])
    worst = 0.0; flips = 0
    nchk = min(r["steps"], 40)
    for k in range(nchk):
        go = s.step(pk, vk, ak, l)
        oo = orc.step(O, pk, vk, ak, cfg["pf"], l, cfg["pmin"], cfg["pmax"], nthreads=8)
        same = (go["status"] & 0xff) == (oo["status"] & 0xff)
        flips += int((~same).sum())
        d = np.abs(go["l_new"] - oo["l_new"]).max(axis=(0, 1))
        worst = max(worst, d[same].max())
        # device-resident trajectory must equal the host-stepped one
        dres = np.abs(r["pk"][:, k + 1, :] - go["p1"]).max()
        if dres > 1e-9: print("  step", k, "resident vs host-stepped", dres)
        l, pk, vk, ak = go["l_new"], go["p1"], go["v1"], go["a1"]
    print(name, "teacher-forced %d steps: max |GPU-oracle| %.3e, status flips %d" % (nchk, worst, flips))
    it = r["status_hist"]
    print(name, "status histogram", {int(k): int(v) for k, v in zip(*np.unique(it & 0xff, return_counts=True))})
    for mode in (0, 1, 2):
        s.init_horizons(cfg["po"])
        r2 = s.run(cfg["max_steps"], mode=mode)
        tm = s.last_timing()
        print(name, "mode", mode, "steps", r2["steps"], "ms/step %.4f scan %.4f qp %.4f -> %.3e agent-steps/s" % (
            tm["step_ms"], tm["scan_ms"], tm["qp_ms"], N / (tm["step_ms"] * 1e-3)))
    s.close()
