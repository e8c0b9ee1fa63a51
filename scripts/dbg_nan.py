import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import dmpc, scenarios
cfg = scenarios.config("C4")
P = dmpc.default_params(cfg["variant"], **cfg["params"])
N = cfg["N"]
with dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
    print(s.config())
    l, pk, vk, ak = s.init_horizons(cfg["po"])
    for k in range(6):
        g = s.step(pk, vk, ak, l, want_horizons=True)
        bad = np.nonzero(np.isnan(g["l_new"]).any(axis=(0, 1)) | np.isnan(g["a_hor"]).any(axis=(0, 1)))[0]
        print(k, "nan agents", bad[:10], "status", np.unique(g["status"] & 0xff, return_counts=True), "it max", g["diag"]["iters"].max(),
              "nv max", g["diag"]["nv"].max(), "nact max", g["diag"]["nact"].max())
        if len(bad):
            print("diag of bad", g["diag"][bad[:10]], g["status"][bad[:10]])
            np.savez("gpurun_out/nan_case.npz", pk=pk, vk=vk, ak=ak, l=l, bad=bad, l_new=g["l_new"], status=g["status"])
            break
        l, pk, vk, ak = g["l_new"], g["p1"], g["v1"], g["a1"]
