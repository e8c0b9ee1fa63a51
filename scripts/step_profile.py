"""Per-step timing over a whole closed-loop transition (plain launches, CUDA events per kernel)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import dmpc, scenarios
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = scenarios.config(name)
P = dmpc.default_params(cfg["variant"], **cfg["params"])
N = cfg["N"]
with dmpc.Solver(N, P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
    for rep in range(2):
        s.init_horizons(cfg["po"])
        rows = []
        for k in range(cfg["max_steps"]):
            r = s.run(1, mode=1)
            t = s.last_timing()
            st = s.get_state()
            it = st["diag"]["iters"]
            tries = (st["status"] >> 8) & 0xff
            rows.append((k, t["scan_ms"] * 1e3, t["qp_ms"] * 1e3, t["step_ms"] * 1e3, it.mean(), it.max(), int(np.sort(it)[-5]),
                         int(tries.max()), int((tries > 0).sum()), int(((st["status"] & 1) == 0).sum())))
            if r["reached"]:
                break
    a = np.array(rows)
    print("step  scan_us   qp_us  step_us  it_mean it_max it_top5 tries_max n_retry n_fail")
    for r in rows:
        if r[0] < 30 or r[0] % 10 == 0:
            print("%4d %8.1f %8.1f %8.1f %7.1f %6d %6d %6d %6d %6d" % r)
    print("mean scan %.1f us, qp %.1f us, step %.1f us -> %.3e agent-steps/s" % (a[:, 1].mean(), a[:, 2].mean(), a[:, 3].mean(),
                                                                       N / (a[:, 3].mean() * 1e-6)))
