"""Scenario generation (host side, like the reference): randomTest.m restated with numpy's RNG.

Reference: dmpc/matlab/randomTest.m:1-57 (rejection sampling of start and goal sets with pairwise
ellipsoid distance > rmin), test/failure_rate.m:63-64 (arena from the 1 agent/m^3 density rule).
MATLAB's RNG stream cannot be reproduced, so scenarios are seeded numpy draws; the same arrays are
handed to the CPU baseline and to the GPU path.
"""
from __future__ import annotations

import numpy as np


def density_arena(N: int, density: float = 1.0):
    """failure_rate.m:63-64: pmin = [-s/2,-s/2,0.2], pmax = [s/2,s/2,s+0.2], s = (N/density)^(1/3)."""
    s = (N / density) ** (1.0 / 3.0)
    return np.array([-s / 2, -s / 2, 0.2]), np.array([s / 2, s / 2, s + 0.2])


def _sample_set(N, pmin, pmax, rmin, c, rng, max_tries=100000):
    E1 = np.array([1.0, 1.0, 1.0 / c])
    pts = np.zeros((3, N))
    n = 0
    tries = 0
    while n < N:
        cand = pmin + (pmax - pmin) * rng.random(3)
        tries += 1
        if n == 0 or (np.sqrt((((pts[:, :n] - cand[:, None]) * E1[:, None]) ** 2).sum(0)) > rmin).all():
            pts[:, n] = cand
            n += 1
            tries = 0
        elif tries > max_tries:
            n, tries = 0, 0  # start over like randomTest.m does when it cannot place a point
    return pts


def random_test(N, pmin, pmax, rmin, c, seed):
    """Returns po, pf as (3, N) float64."""
    rng = np.random.default_rng(seed)
    pmin, pmax = np.asarray(pmin, float), np.asarray(pmax, float)
    return _sample_set(N, pmin, pmax, rmin, c, rng), _sample_set(N, pmin, pmax, rmin, c, rng)


def config(name: str):
    """The workloads of BASELINE.json / SURVEY.md section 8(d): dict(N, K, variant, pmin, pmax, seed, ...)."""
    from . import dmpc
    if name == "C1":  # 4-agent corner swap, dmpc_soft_bound.m:43-54
        po = np.array([[1.501, 1.5, 1.5], [-1.5, -1.5, 1.5], [-1.5, 1.5, 1.5], [1.5, -1.5, 1.5]]).T
        pf = np.array([[-1.5, -1.5, 1.5], [1.5, 1.5, 1.5], [1.5, -1.5, 1.5], [-1.5, 1.5, 1.5]]).T
        return dict(N=4, K=15, variant=dmpc.SOFT_BOUND, pmin=np.array([-2.5, -2.5, 0.2]),
                    pmax=np.array([2.5, 2.5, 2.2]), po=po, pf=pf, params=dict(rmin=0.5, c=1.5), max_steps=100)
    if name == "C2":
        pmin, pmax = np.array([-5, -5, 0.2]), np.array([5, 5, 3.2])
        po, pf = random_test(100, pmin, pmax, 0.35, 2.0, 1002)
        return dict(N=100, K=15, variant=dmpc.HARD, pmin=pmin, pmax=pmax, po=po, pf=pf, params=dict(),
                    max_steps=149)
    if name in ("C3", "N100", "N500", "N2000"):
        N = {"C3": 500, "N100": 100, "N500": 500, "N2000": 2000}[name]
        pmin, pmax = density_arena(N)
        po, pf = random_test(N, pmin, pmax, 0.35, 2.0, 1003)
        return dict(N=N, K=15, variant=dmpc.SOFT_BOUND, pmin=pmin, pmax=pmax, po=po, pf=pf, params=dict(),
                    max_steps=149)
    if name == "C4":
        pmin, pmax = np.array([-5, -5, 0.2]), np.array([5, 5, 10.2])
        po, pf = random_test(2000, pmin, pmax, 0.35, 2.0, 1004)
        return dict(N=2000, K=20, variant=dmpc.SOFT_BOUND, pmin=pmin, pmax=pmax, po=po, pf=pf, params=dict(K=20),
                    max_steps=149)
    if name == "C5":
        # 100 Monte-Carlo trials x N = 200 (test/failure_rate.m:61-68 shape, 1 agent/m^3 arena), seeds 2000..2099:
        # po / pf are lists, one (3, N) pair per scenario; batched by Solver(n_scenarios=100)
        N, S = 200, 100
        pmin, pmax = density_arena(N)
        pairs = [random_test(N, pmin, pmax, 0.35, 2.0, 2000 + s) for s in range(S)]
        return dict(N=N, K=15, S=S, variant=dmpc.SOFT_BOUND, pmin=pmin, pmax=pmax, po=[p[0] for p in pairs],
                    pf=[p[1] for p in pairs], params=dict(), max_steps=149)
    raise KeyError(name)
