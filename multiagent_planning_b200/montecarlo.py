"""Monte-Carlo experiment harness in the shape of the reference's test/failure_rate.m (and, with two variants
on identical scenarios, test/comp_kctr.m / comp_hardsoft2.m): for every swarm size N a batch of random trials at
constant density, every trial a complete DMPC transition, every outcome classified like the script does.

    test/failure_rate.m:56-68     arena from N (1 agent / m^3), [po, pf] = randomTest(...) per trial
                 :99-127          the MPC loop of a trial; it ends at the goal, at max_K, or when a QP is
                                  infeasible (`coll` / `outbound` do NOT end it: feasible stays 1)
                 :128-195         goal check, time scaling, 100 Hz splines, pairwise post-check, statistics
                 :196, :253-258   success = feasible & ~failed_goal & ~violation; failure taxonomy

Here ALL trials of a swarm size are one batched handle: scenario generation on the device
(dmpcb200_gen_scenarios), one dmpcb200_run_batch for the loops (three launches per MPC step for the whole
batch), post-processing per finished trial on the device.  The random scenarios are not MATLAB's (its stream is
unseeded); the statistics are the comparable quantity (see `published()`).
"""
from __future__ import annotations

import numpy as np

from . import dmpc, scenarios


def failure_rate(N_vector=(20, 40, 60, 80, 100, 120, 140, 160, 180, 200), trials=50, seed=0, variant=dmpc.SOFT_BOUND,
                 max_T=30.0, rmin_init=0.35, device=0, mode=0, postprocess=True, **params):
    """Returns a dict of (len(N_vector), trials) arrays with the reference workspace's names -- feasible,
    failed_goal, violation, success_dmpc, traj_time, totdist_dmpc, steps -- plus prob_dmpc, t_dmpc (device
    seconds per trial: loop + post-processing, the reference's tic/toc region) and the failure taxonomy."""
    P = dmpc.default_params(variant, **params)
    nq = len(N_vector)
    max_K = int(round(max_T / P.h)) + 1
    out = {k: np.zeros((nq, trials)) for k in ("feasible", "failed_goal", "violation", "success_dmpc", "steps")}
    for k in ("traj_time", "totdist_dmpc", "t_dmpc", "min_dist"):
        out[k] = np.full((nq, trials), np.nan)
    for q, N in enumerate(N_vector):
        pmin, pmax = scenarios.density_arena(int(N))                      # failure_rate.m:63-64
        with dmpc.Solver(int(N), P, n_scenarios=trials, device=device) as b:
            b.gen_scenarios(seed + 1000 * q, pmin, pmax, rmin_init=rmin_init, mode=mode, want_points=False)
            r = b.run_batch(max_K - 2, stop_on_fail=2, record=True)
            loop_s = r["device_ms"] * 1e-3 / trials
            reached = np.asarray(r["reached"], bool)
            infeasible = (~reached) & (r["steps"] < max_K - 2)            # the loop ended early without the goal
            out["feasible"][q] = ~infeasible
            out["failed_goal"][q] = (~infeasible) & (~reached)
            out["steps"][q] = r["steps"]
            for t in range(trials):
                if infeasible[t] or not reached[t]:
                    continue
                out["t_dmpc"][q, t] = loop_s
                if not postprocess or r["steps"][t] < 3:
                    continue
                pp = b.postprocess(r["pk"][t], r["vk"][t], r["ak"][t], want_interp=False, scenario=t)
                out["violation"][q, t] = pp["violation"]
                out["traj_time"][q, t] = pp["traj_time"]
                out["totdist_dmpc"][q, t] = pp["totdist"]
                out["min_dist"][q, t] = pp["min_dist"]
                out["t_dmpc"][q, t] = loop_s + pp["device_ms"] * 1e-3
        out["success_dmpc"][q] = (out["feasible"][q] > 0) & (out["failed_goal"][q] == 0) & (out["violation"][q] == 0)
    out["N_vector"] = np.asarray(N_vector)
    out["prob_dmpc"] = out["success_dmpc"].sum(1) / trials                 # failure_rate.m:222
    infes, viol, goal = (1 - out["feasible"]).sum(1), out["violation"].sum(1), out["failed_goal"].sum(1)
    out["taxonomy"] = dict(infeasible=infes, collisions=viol, incomplete=goal)   # :253-258
    return out


def published():
    """the reference's own curve (data/failure_rate/failure_rate3.mat, 50 trials per N, MATLAB quadprog)"""
    return dict(N_vector=np.arange(20, 201, 20),
                prob_dmpc=np.array([1, 1, 1, .94, .90, .84, .78, .64, .56, .28]),
                t_dmpc_mean_s=np.array([6.4, 15.2, 26.7, 42.0, 62.2, 82.6, 108.7, 136.5, 164.9, 195.4]))


def compare_variants(N, trials, variants=(dmpc.SOFT_BOUND, dmpc.SOFT_BOUND2), seed=0, **kw):
    """test/comp_kctr.m / comp_hardsoft2.m shape: several solver variants on IDENTICAL random scenarios"""
    return {int(v): failure_rate((N,), trials, seed=seed, variant=v, **kw) for v in variants}
