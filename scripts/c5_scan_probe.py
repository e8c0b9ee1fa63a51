import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import dmpc, scenarios
S = 100
cfg = scenarios.config("C5")
P = dmpc.default_params(0)
with dmpc.Solver(cfg["N"], P, n_scenarios=S) as b:
    for rep in range(2):
        for s in range(S):
            b.set_scenario(s, cfg["po"][s], cfg["pf"][s], cfg["pmin"], cfg["pmax"])
        r = b.run_batch(30, stop_on_fail=False, mode=1)
    print(os.environ.get("DMPCB200_SCAN_LAYOUT"), os.environ.get("DMPCB200_LAYOUT"), "first 30 steps:", {k: round(1e3*v,1) for k,v in b.last_timing().items() if k != 'launches'})
    r = b.run_batch(119, stop_on_fail=False, mode=1)
    print("   next 119 steps:", {k: round(1e3*v,1) for k,v in b.last_timing().items() if k != 'launches'})
