"""Per-agent solve time of the QP kernel (needs a -DDMPC_PROF_AGENT build as libdmpc_b200_prof.so: the
diag word `nact` then carries cycles/16).  Prints the slowest agents of a few steps."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libdmpc_b200_prof.so")
from multiagent_planning_b200 import dmpc, scenarios
cfg = scenarios.config(sys.argv[1] if len(sys.argv) > 1 else "C3")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
P = dmpc.default_params(cfg["variant"], **cfg["params"])
with dmpc.Solver(cfg["N"], P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
    s.init_horizons(cfg["po"])
    if skip:
        s.run(skip, mode=2)
    for k in range(skip, skip + steps):
        s.run(1, mode=1)
        t = s.last_timing()
        st = s.get_state()
        d = st["diag"]
        us = d["nact"].astype(np.float64) * 16 / 1965.0
        tries = (st["status"] >> 8) & 0xff
        order = np.argsort(-us)[:6]
        print("step %d qp %.1f us | slowest agents: " % (k, t["qp_ms"] * 1e3) +
              "; ".join("n%d %.0fus it%d nv%d tr%d k*%d" % (n, us[n], d["iters"][n], d["nv"][n], tries[n], d["kstar"][n]) for n in order)
              + " | median %.1f us, agents with 0 it: %.1f us" % (np.median(us), np.median(us[d["iters"] == 0]) if (d["iters"] == 0).any() else -1))
