"""CPU probe (host build of the device algorithm): iterations of the slowest agent per step, cold vs warm start.
Run as  EMUL_FLAGS="-DDMPC_WARM_START -DDMPC_WARM_STATS [-DDMPC_WARM_NOROWS] [-DDMPC_WARM_SHIFT=0]" python scripts/warm_probe.py C3 26
(the experiment is not part of the product build; touch tests/host_emul/emul.cpp afterwards to rebuild the plain test library)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import scenarios
from oracle import dmpc_oracle as orc
from tests.host_emul import emul
emul.build(force=True)
orc.build()
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 26
cfg = scenarios.config(name)
N = cfg["N"]
P = orc.default_params(cfg["variant"])
for k, v in cfg["params"].items():
    setattr(P, k, v)
K = P.K
po, pf, pmin, pmax = cfg["po"], cfg["pf"], cfg["pmin"], cfg["pmax"]
l = np.zeros((3, K, N), order="F")
for n in range(N):
    l[:, :, n] = orc.init_dmpc(po[:, n], pf[:, n], P.h, K, P.init_div)[0]
pk, vk, ak = l[:, 0, :].copy(), np.zeros((3, N)), np.zeros((3, N))
warm = np.zeros((N, 68), np.int32)
EP = emul.params_from(P)
tot = [0, 0]
for k in range(steps):
    c = emul.step(EP, pk, vk, ak, pf, l, pmin, pmax, QMAX=-64)
    w = emul.step(EP, pk, vk, ak, pf, l, pmin, pmax, QMAX=-64, warm=warm)
    assert np.array_equal(c["status"], w["status"]), (k, np.nonzero(c["status"] != w["status"]))
    err = np.abs(c["l_new"] - w["l_new"]).max()
    ci, wi = c["diag"][:, 2], w["diag"][:, 2]
    wq = (w["diag"][:, 3] >> 8) & 0xff      # active set after the warm start
    wb = (w["diag"][:, 3] >> 16) & 0xff     # assembled
    wl = (w["diag"][:, 3] >> 24) & 0xff     # list
    wn = w["diag"][:, 3] & 0xff
    cost_w = wi + 0.4 * wb           # bordering ~ 0.4 iteration per constraint
    ic, iw = int(ci.argmax()), int(cost_w.argmax())
    print(f"step {k:3d} cold max {ci.max():4d} (sum {ci.sum():6d})  warm iters max {wi.max():4d} cost max {cost_w.max():6.1f} "
          f"[agent {iw}: iters {wi[iw]} list {wl[iw]} built {wb[iw]} q0 {wq[iw]} q {wn[iw]} tries {(w['status'][iw]>>8)&255} nv {w['diag'][iw,1]}; cold {ci[iw]}] sum {wi.sum():6d}+{0.4*wb.sum():.0f}  err {err:.1e}")
    if k >= 5:
        tot[0] += ci.max(); tot[1] += cost_w.max()
    l, pk, vk, ak = c["l_new"], c["p1"], c["v1"], c["a1"]
print("slowest-agent cost, steps 5..: cold", tot[0], "warm", tot[1])
