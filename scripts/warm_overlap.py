"""Overlap of the final QP active sets of consecutive MPC steps (heaviest agents), host build with
EMUL_FLAGS="-DDMPC_WARM_START" (the experiment build stores the final set); see profiles/r2c_warm_start_probe.txt."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import scenarios
from oracle import dmpc_oracle as orc
from tests.host_emul import emul
emul.build(force=True); orc.build()
cfg = scenarios.config("C3"); N = cfg["N"]
P = orc.default_params(cfg["variant"])
for k, v in cfg["params"].items(): setattr(P, k, v)
K = P.K
po, pf, pmin, pmax = cfg["po"], cfg["pf"], cfg["pmin"], cfg["pmax"]
l = np.zeros((3, K, N), order="F")
for n in range(N): l[:, :, n] = orc.init_dmpc(po[:, n], pf[:, n], P.h, K, P.init_div)[0]
pk, vk, ak = l[:, 0, :].copy(), np.zeros((3, N)), np.zeros((3, N))
EP = emul.params_from(P)
prev = None
def sets(w):
    out = []
    for n in range(N):
        cnt = w[n, 0]
        out.append(set(int(c) for c in w[n, 1:1 + cnt]))
    return out
names = {0:"BL",1:"BU",2:"WL",3:"WU",4:"ROW"}
def fmt(c):
    t=(c>>16)&0xff; i=c&0xffff
    if t<4: return f"{names[t]}{i//3}{'xyz'[i%3]}"
    return f"R{i}{'u' if c&(1<<24) else ''}{'l' if c&(1<<25) else ''}"
for k in range(10):
    warm = np.zeros((N, 68), np.int32)   # fresh: cold solve, but the final set is stored
    c = emul.step(EP, pk, vk, ak, pf, l, pmin, pmax, QMAX=-64, warm=warm)
    cur = sets(warm)
    if prev is not None:
        it = c["diag"][:, 2]
        heavy = np.argsort(-it)[:6]
        for n in heavy:
            pred = prev[n]
            actual = set((x + 3) if ((x >> 16) & 0xff) < 4 else x for x in cur[n])   # un-shift: index in THIS step
            pred_noshift = set((x + 3) if ((x >> 16) & 0xff) < 4 else x for x in pred)
            print(f"step {k} agent {n} iters {it[n]} |actual| {len(actual)} shift: hit {len(pred & actual)} wrong {len(pred - actual)} | noshift: hit {len(pred_noshift & actual)} wrong {len(pred_noshift - actual)}  kstar {c['diag'][n,0]} nv {c['diag'][n,1]}")
            if k in (3,4) and n == heavy[0]:
                print("   pred  ", sorted(fmt(x) for x in pred))
                print("   actual", sorted(fmt(x) for x in actual))
    prev = cur
    l, pk, vk, ak = c["l_new"], c["p1"], c["v1"], c["a1"]
