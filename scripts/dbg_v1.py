import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiagent_planning_b200 import dmpc, scenarios
from oracle import dmpc_oracle as orc
cfg = scenarios.config("C3"); N = cfg["N"]
po, pf = scenarios.random_test(N, cfg["pmin"], cfg["pmax"], 0.35, 2.0, 11)
P = dmpc.default_params(1)
O = orc.default_params(1)
with dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=pf) as s:
    l, pk, vk, ak = s.init_horizons(po)
    for rep in range(1):
        g = s.step(pk, vk, ak, l)
        if rep == 0:
            o = orc.step(O, pk, vk, ak, pf, l, cfg["pmin"], cfg["pmax"], nthreads=8)
        bad = np.nonzero((g["status"] & 0xFFFF) != (o["status"] & 0xFFFF))[0]
        print("rep", rep, "mismatching agents", bad, [hex(x) for x in g["status"][bad]], [hex(x) for x in o["status"][bad]],
              "diag", [tuple(g["diag"][b]) for b in bad], "err", np.abs(g["l_new"] - o["l_new"]).max())
    n = 160
    print("agent 160: gpu %#x oracle %#x diag %s" % (g["status"][n], o["status"][n], tuple(g["diag"][n])))
