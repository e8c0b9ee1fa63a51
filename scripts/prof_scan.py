"""Phase timeline of the scan kernel (needs a -DDMPC_PROF_SCAN build:
cd multiagent_planning_b200/csrc && nvcc ... -DDMPC_PROF_SCAN -o ../libdmpc_b200_prof.so ...).
Prints, per phase end, the max and mean over warps of the cycles since kernel entry."""
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
names = ["setup (barriers, TMA issue, own)", "tile loop", "combine barrier", "scan_finish (rows)", "of which mbarrier wait"]
with dmpc.Solver(cfg["N"], P, pmin=cfg["pmin"], pmax=cfg["pmax"], pf=cfg["pf"]) as s:
    s.init_horizons(cfg["po"])
    out = (C.c_uint64 * 32)()
    s.run(2, mode=2)
    L.dmpcb200_prof_read(out)
    for k in range(steps):
        s.run(1, mode=2)
        L.dmpcb200_prof_read(out)
        o = np.array(out[:], dtype=np.float64)
        print("step %2d  " % k + "  ".join("%s: max %6.0f mean %6.0f" % (names[i].split(" (")[0], o[i], o[8 + i] / max(o[16 + i], 1))
                                          for i in range(5)), " timing", s.last_timing()["scan_ms"])
