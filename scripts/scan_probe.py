import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import dmpc, scenarios
for name in sys.argv[1:]:
    cfg = scenarios.config(name)
    P = dmpc.default_params(cfg["variant"], **cfg["params"])
    with dmpc.Solver(cfg["N"], P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
        for rep in range(2):
            s.init_horizons(cfg["po"])
            r = s.run(30, mode=1)
        print(name, os.environ.get("DMPCB200_SCAN_PRUNE"), {k: round(1e3 * v, 1) for k, v in s.last_timing().items() if k != "launches"})
