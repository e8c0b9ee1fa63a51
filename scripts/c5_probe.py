"""C5 probe: 100 scenarios x N=200 batched in one handle; device time, agent-steps/s, outcome statistics."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiagent_planning_b200 import dmpc, scenarios

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 149
cfg = scenarios.config("C5")
P = dmpc.default_params(0)
with dmpc.Solver(cfg["N"], P, n_scenarios=S) as b:
    for rep in range(2):
        for s in range(S):
            b.set_scenario(s, cfg["po"][s], cfg["pf"][s], cfg["pmin"], cfg["pmax"])
        t0 = time.perf_counter()
        r = b.run_batch(steps, stop_on_fail=True)
        dt = time.perf_counter() - t0
        print(f"rep {rep}: device {r['device_ms']:.2f} ms wall {1e3*dt:.2f} ms, agent-steps {r['agent_steps']}, "
              f"{r['agent_steps']/(r['device_ms']*1e-3)/1e6:.2f} M agent-steps/s; reached {int(r['reached'].sum())}/{S}, "
              f"failed {(r['first_fail_step']>=0).sum()}, steps mean {r['steps'].mean():.1f} max {r['steps'].max()}")
    for s in range(S):
        b.set_scenario(s, cfg["po"][s], cfg["pf"][s], cfg["pmin"], cfg["pmax"])
    r = b.run_batch(steps, stop_on_fail=False)
    print(f"no stop: device {r['device_ms']:.2f} ms, {r['agent_steps']/(r['device_ms']*1e-3)/1e6:.2f} M agent-steps/s; reached {int(r['reached'].sum())}/{S} steps mean {r['steps'].mean():.1f}")
    for s in range(S):
        b.set_scenario(s, cfg["po"][s], cfg["pf"][s], cfg["pmin"], cfg["pmax"])
    r = b.run_batch(30, stop_on_fail=False, mode=1)
    print("first 30 steps, per-kernel:", b.last_timing(), r["device_ms"])
