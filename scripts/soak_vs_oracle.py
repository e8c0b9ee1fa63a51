"""Soak check (not part of the test-suite): a whole closed-loop transition on the GPU, every step teacher-forced
against the CPU oracle: status flags, retry counts and horizons.
usage: soak_vs_oracle.py <workload>[:seed[:variant]] [steps]
  workload: C2 | C3 | N100 | N2000 | C4 | C5 (= every scenario of the batch, one after the other)
  seed: other start / goal sets for the same configuration; variant: 0..3 (overrides the configuration's)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import dmpc, scenarios
from oracle import dmpc_oracle as orc
spec = (sys.argv[1] if len(sys.argv) > 1 else "C3").split(":")
name = spec[0]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 149
cfg = scenarios.config(name)
variant = int(spec[2]) if len(spec) > 2 and spec[2] != "" else cfg["variant"]
P = dmpc.default_params(variant, **cfg["params"])
O = orc.default_params(variant)
for k, v in cfg["params"].items():
    setattr(O, k, v)
N = cfg["N"]
if name == "C5":
    cases = [(f"C5[{s}]", cfg["po"][s], cfg["pf"][s]) for s in range(cfg["S"])]
elif len(spec) > 1 and spec[1] != "":
    po, pf = scenarios.random_test(N, cfg["pmin"], cfg["pmax"], 0.35, 2.0, int(spec[1]))
    cases = [(sys.argv[1], po, pf)]
else:
    cases = [(name, cfg["po"], cfg["pf"])]
tot_steps = tot_retr = tot_flips = 0
tot_worst = 0.0
with dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cases[0][2]) as s:
    for label, po, pf in cases:
        s.set_goals(pf)
        worst, flips, retr = 0.0, 0, 0
        l, pk, vk, ak = s.init_horizons(po)
        for k in range(steps):
            g = s.step(pk, vk, ak, l)
            o = orc.step(O, pk, vk, ak, pf, l, cfg["pmin"], cfg["pmax"], nthreads=os.cpu_count() or 1)
            same = np.array_equal(g["status"] & 0xFF, o["status"] & 0xFF)
            tr_same = np.array_equal((g["status"] >> 8) & 0xFF, (o["status"] >> 8) & 0xFF)
            err = float(np.abs(g["l_new"] - o["l_new"]).max())
            worst = max(worst, err)
            retr += int((((o["status"] >> 8) & 0xFF) > 0).sum())
            if not same or not tr_same or err > 1e-6:
                flips += 1
                bad = np.nonzero((g["status"] & 0xFFFF) != (o["status"] & 0xFFFF))[0]
                print(f"{label} step {k}: status equal {same}, retries equal {tr_same}, max |dl| {err:.2e}, agents {bad[:8]}")
            l, pk, vk, ak = o["l_new"], o["p1"], o["v1"], o["a1"]
            if orc.reached_goal(pk, pf, 0.01)[0]:
                break
        if len(cases) == 1:
            print(f"{label}: {k + 1} steps, {retr} retried agent-steps, mismatching steps {flips}, worst |GPU - oracle| = {worst:.2e} m")
        tot_steps += k + 1; tot_retr += retr; tot_flips += flips; tot_worst = max(tot_worst, worst)
if len(cases) > 1:
    print(f"{name}: {len(cases)} scenarios, {tot_steps} steps, {tot_retr} retried agent-steps, mismatching steps {tot_flips}, "
          f"worst |GPU - oracle| = {tot_worst:.2e} m")
