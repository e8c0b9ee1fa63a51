import sys, os, subprocess, pickle
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if len(sys.argv) > 1:
    from multiagent_planning_b200 import dmpc, scenarios
    N = 900
    pmin, pmax = scenarios.density_arena(N, density=1.5)
    po, pf = scenarios.random_test(N, pmin, pmax, 0.35, 2.0, seed=99)
    P = dmpc.default_params(0)
    with dmpc.Solver(N, P, pmin=pmin, pmax=pmax, pf=pf) as s:
        s.init_horizons(po)
        out = []
        for k in range(14):
            r = s.run(1, record=True, status_hist=True)
            st = s.get_state()
            out.append((st["pk"].copy(), st["status"].copy(), st["diag"].copy(), st["l"].copy()))
    pickle.dump(out, open(sys.argv[1], "wb"))
else:
    for lay, f in (("auto", "/tmp/a.pkl"), ("classic", "/tmp/c.pkl")):
        env = dict(os.environ)
        if lay == "classic": env["DMPCB200_LAYOUT"] = "classic"
        subprocess.check_call([sys.executable, __file__, f], env=env)
    a, c = pickle.load(open("/tmp/a.pkl", "rb")), pickle.load(open("/tmp/c.pkl", "rb"))
    for k in range(14):
        d = np.abs(a[k][3] - c[k][3]).max(axis=(0, 1))
        bad = np.nonzero(d > 0)[0]
        print("step", k, "differing agents", len(bad), "max diff %.3e" % d.max(), "status auto/classic", [(int(i), hex(int(a[k][1][i])), hex(int(c[k][1][i])), tuple(a[k][2][i]), tuple(c[k][2][i])) for i in bad[:5]])
        if len(bad): break
