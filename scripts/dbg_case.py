"""debug: one teacher-forced step of <workload>:<seed>:<variant> at <step>, GPU vs oracle, mismatching agents"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiagent_planning_b200 import dmpc, scenarios
from oracle import dmpc_oracle as orc
name, seed, variant, step = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = scenarios.config(name); N = cfg["N"]
po, pf = scenarios.random_test(N, cfg["pmin"], cfg["pmax"], 0.35, 2.0, seed)
P = dmpc.default_params(variant, **cfg["params"])
O = orc.default_params(variant)
for k, v in cfg["params"].items(): setattr(O, k, v)
with dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=pf) as s:
    l, pk, vk, ak = s.init_horizons(po)
    for k in range(step + 1):
        o = orc.step(O, pk, vk, ak, pf, l, cfg["pmin"], cfg["pmax"], nthreads=os.cpu_count())
        if k == step:
            g = s.step(pk, vk, ak, l)
            bad = np.nonzero((g["status"] & 0xFFFF) != (o["status"] & 0xFFFF))[0]
            print("mismatching agents", bad, [hex(x) for x in g["status"][bad]], [hex(x) for x in o["status"][bad]],
                  "diag", [tuple(int(v) for v in g["diag"][b]) for b in bad], "err", np.abs(g["l_new"] - o["l_new"]).max())
        l, pk, vk, ak = o["l_new"], o["p1"], o["v1"], o["a1"]
