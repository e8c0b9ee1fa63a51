"""Post-processing of the C3 transition (dmpcb200_postprocess) for profiling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import dmpc, scenarios
cfg = scenarios.config(sys.argv[1] if len(sys.argv) > 1 else "C3")
P = dmpc.default_params(cfg["variant"], **cfg["params"])
with dmpc.Solver(cfg["N"], P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
    s.init_horizons(cfg["po"])
    tr = s.run(cfg["max_steps"], record=True)
    for _ in range(3):
        pp = s.postprocess(tr["pk"], tr["vk"], tr["ak"], want_interp=False)
    print({k: pp[k] for k in ("r_factor", "nt", "min_dist", "violation", "totdist", "traj_time", "device_ms")})
