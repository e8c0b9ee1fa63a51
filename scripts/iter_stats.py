"""Per-agent QP iteration counts over closed-loop steps (saved for offline study of the persistent grid's queue order)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import dmpc, scenarios
out = {}
for name in sys.argv[1:]:
    cfg = scenarios.config(name)
    P = dmpc.default_params(cfg["variant"], **cfg["params"])
    N = cfg["N"]
    with dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        s.init_horizons(cfg["po"])
        its, nvs, qs, tq = [], [], [], []
        for k in range(26):
            s.run(1, mode=1)
            t = s.last_timing()
            st = s.get_state()
            its.append(st["diag"]["iters"].copy()); nvs.append(st["diag"]["nv"].copy()); qs.append(st["diag"]["nact"].copy())
            tq.append(t["qp_ms"] * 1e3)
        out[name + "_iters"] = np.array(its); out[name + "_nv"] = np.array(nvs); out[name + "_nact"] = np.array(qs)
        out[name + "_qp_us"] = np.array(tq)
        print(name, "qp_us", np.round(tq, 1))
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/iter_stats.npz", **out)
