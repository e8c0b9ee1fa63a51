"""Short profiling driver: N=500 C3 workload, a few closed-loop steps through dmpcb200_run (plain launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import dmpc, scenarios
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cfg = scenarios.config(name)
P = dmpc.default_params(cfg["variant"], **cfg["params"])
with dmpc.Solver(cfg["N"], P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
    s.init_horizons(cfg["po"])
    r = s.run(steps, mode=2)
    print(name, r, s.last_timing())
