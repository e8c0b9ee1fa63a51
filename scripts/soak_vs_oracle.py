"""Soak check (not part of the test-suite): a whole closed-loop transition on the GPU, every step teacher-forced
against the CPU oracle: status flags, retry counts and horizons.  usage: soak_vs_oracle.py [C3|N100|C4|...] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import dmpc, scenarios
from oracle import dmpc_oracle as orc
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 149
cfg = scenarios.config(name)
P = dmpc.default_params(cfg["variant"], **cfg["params"])
O = orc.default_params(cfg["variant"])
for k, v in cfg["params"].items():
    setattr(O, k, v)
N = cfg["N"]
worst, flips, retr = 0.0, 0, 0
with dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
    l, pk, vk, ak = s.init_horizons(cfg["po"])
    for k in range(steps):
        g = s.step(pk, vk, ak, l)
        o = orc.step(O, pk, vk, ak, cfg["pf"], l, cfg["pmin"], cfg["pmax"], nthreads=os.cpu_count() or 1)
        same = np.array_equal(g["status"] & 0xFF, o["status"] & 0xFF)
        tr_same = np.array_equal((g["status"] >> 8) & 0xFF, (o["status"] >> 8) & 0xFF)
        err = float(np.abs(g["l_new"] - o["l_new"]).max())
        worst = max(worst, err)
        retr += int((((o["status"] >> 8) & 0xFF) > 0).sum())
        if not same or not tr_same or err > 1e-6:
            flips += 1
            print(f"step {k}: status equal {same}, retries equal {tr_same}, max |dl| {err:.2e}")
        l, pk, vk, ak = o["l_new"], o["p1"], o["v1"], o["a1"]
        if orc.reached_goal(pk, cfg["pf"], 0.01)[0]:
            break
print(f"{name}: {k + 1} steps, {retr} retried agent-steps, mismatching steps {flips}, worst |GPU - oracle| = {worst:.2e} m")
