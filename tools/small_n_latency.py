"""Development aid: wall time per setulb call at small n (launch/latency-bound regime)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import harness as H
from oracle import oracle_py as O
import lbfgsb_b200

for n, m in ((1000, 10), (100000, 10)):
    x, l, u, nbd = H.rosenbrock_problem(n, l_odd=1.1)
    s = lbfgsb_b200.HostSetulb()
    t0 = time.perf_counter()
    tr = H.run_driver(s, O.rosenbrock_fg, n, m, x, l, u, nbd, 0.0, 0.0, stop=H.iteration_budget_stop(40), want_hash=False)
    dt = time.perf_counter() - t0
    print("n=%d m=%d: %d iterations, %d fg, %.1f ms total, %.3f ms per iteration" % (n, m, len(tr[0]), tr[0][-1]["nfgv"], dt * 1e3, dt * 1e3 / len(tr[0])))

import torch
for n, m in ((1000, 10), (3000, 10), (100000, 10)):
    x, l, u, nbd = H.rosenbrock_problem(n, l_odd=1.1)
    xd, ld, ud, nd = (torch.from_numpy(a).cuda() for a in (x, l, u, nbd))
    gd = torch.zeros_like(xd)
    prob = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    fg = lbfgsb_b200.RosenbrockDevice(np.float64)
    prob.profile(True)
    while True:
        prob.setulb_dev(xd, ld, ud, nd, gd, 0.0, 0.0)
        t = prob.task_str()
        if t[:2] == "FG":
            prob.f[0] = fg(xd, gd)
        elif t[:5] == "NEW_X":
            if prob.isave[29] >= 30:
                break
        else:
            break
    pr = prob.profile_read()
    print("n=%d" % n, {k: (round(v["ms"] / max(v["calls"], 1), 3), v["calls"]) for k, v in pr.items() if v["calls"]})
    prob.close()
