"""CPU soak: the host build of the device algorithm (tests/host_emul, register-resident solver + the kernel's generic
fallback) teacher-forced against the oracle over many seeds / variants -- logic errors of the solver show up here
without a GPU (rounding differs from the device: borderline verdicts are soaked on the GPU, scripts/gpu_soak2.sh).
usage: soak_host_build.py N variant seed0 nseeds steps [density [cpp]]   (cpp: the C++ port's semantics preset of
dmpcb200_default_params_cpp on top of variant 0 / 1: K = 12, neighbour threshold rmin (1 + k/K), slack bound 0.01 doubled for <= 20
retries, term -1e6, rmin 0.5, c 1.5, alim 2)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import scenarios
from oracle import dmpc_oracle as orc
from tests.host_emul import emul
emul.build(); orc.build()
N, variant, seed0, nseeds, steps = (int(x) for x in sys.argv[1:6])
density = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
pmin, pmax = scenarios.density_arena(N, density)
P = orc.default_params(variant)
rmin_init, rmin_goal = 0.35, 2.0
if len(sys.argv) > 7 and sys.argv[7] == "cpp":       # dmpcb200_default_params_cpp (csrc/dmpc_b200.cu), dmpc.cpp:846,907-914,940-945
    assert variant in (0, 1)
    for k_, v_ in dict(neigh_mode=1, slack_lb=-0.01, max_tries=21, term=-1e6, Q1=1000.0, S1=100.0, h=0.2, K=12, c=1.5, rmin=0.5,
                       alim=2.0, goal_tol=0.05, coll_tol=0.05).items():
        setattr(P, k_, v_)
    rmin_init, rmin_goal = 0.5, 1.5
K = P.K
EP = emul.params_from(P)
tot = dict(steps=0, retried=0, bad=0, worst=0.0)
for seed in range(seed0, seed0 + nseeds):
    po, pf = scenarios.random_test(N, pmin, pmax, rmin_init, rmin_goal, seed)
    l = np.zeros((3, K, N), order="F")
    for n in range(N):
        l[:, :, n] = orc.init_dmpc(po[:, n], pf[:, n], P.h, K, P.init_div)[0]
    pk, vk, ak = l[:, 0, :].copy(), np.zeros((3, N)), np.zeros((3, N))
    for k in range(steps):
        o = orc.step(P, pk, vk, ak, pf, l, pmin, pmax, nthreads=8)
        e = emul.step(EP, pk, vk, ak, pf, l, pmin, pmax, QMAX=-64, RMAX=min(N - 1, 256) * (K if variant == 2 else 1))
        same = np.array_equal(o["status"] & 0xFFFF, e["status"] & 0xFFFF)
        err = float(np.abs(o["l_new"] - e["l_new"]).max())
        tot["steps"] += 1
        tot["retried"] += int((((o["status"] >> 8) & 0xFF) > 0).sum())
        tot["worst"] = max(tot["worst"], err)
        if not same or err > 1e-6:
            tot["bad"] += 1
            bad = np.nonzero((o["status"] & 0xFFFF) != (e["status"] & 0xFFFF))[0]
            print(f"MISMATCH N {N} variant {variant} seed {seed} step {k}: agents {bad[:6]} oracle {[hex(x) for x in o['status'][bad[:6]]]} "
                  f"host build {[hex(x) for x in e['status'][bad[:6]]]} err {err:.2e}", flush=True)
        l, pk, vk, ak = o["l_new"], o["p1"], o["v1"], o["a1"]
        if orc.reached_goal(pk, pf, getattr(P, "goal_tol", 0.01))[0]:
            break
print(f"N {N} variant {variant}{' (C++ preset)' if len(sys.argv) > 7 else ''} density {density} seeds {seed0}..{seed0 + nseeds - 1}: {tot['steps']} steps, {tot['retried']} retried agent-steps, "
      f"mismatching steps {tot['bad']}, worst {tot['worst']:.2e} m", flush=True)
