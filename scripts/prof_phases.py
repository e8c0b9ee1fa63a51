"""Per-phase cycle accounting of the QP kernel (needs a -DDMPC_PROF build: scripts/build_prof.sh)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagent_planning_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libdmpc_b200_prof.so")
from multiagent_planning_b200 import dmpc, scenarios
L = _lib.lib()
cfg = scenarios.config(sys.argv[1] if len(sys.argv) > 1 else "C3")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
P = dmpc.default_params(cfg["variant"], **cfg["params"])
names = ["loop-top", "most_violated", "polish", "decode/materialise", "gvec", "mat_vec", "direction: coefs", "direction: apply",
         "direction: zeps + z'Hz", "ratio test", "x,u update+update_P", "border(add)", "drop_slot", "", "", ""]
with dmpc.Solver(cfg["N"], P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
    s.init_horizons(cfg["po"])
    out = (C.c_uint64 * 32)()
    L.dmpcb200_prof_read(out)
    r = s.run(steps, mode=2)
    L.dmpcb200_prof_read(out)
    o = np.array(out[:], dtype=np.float64)
    tot = o[:16].sum()
    print("steps", r["steps"], s.last_timing())
    for i in range(13):
        if o[16 + i]:
            print("%-22s calls %8d  cycles/call %8.0f  share %5.1f%%" % (names[i], o[16 + i], o[i] / o[16 + i], 100 * o[i] / tot))
    st = s.get_state()
    print("iters per agent (last step): mean %.1f max %d" % (st["diag"]["iters"].mean(), st["diag"]["iters"].max()))
