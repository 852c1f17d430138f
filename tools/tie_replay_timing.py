"""Development aid: cost of the heap replay (cauchy_walk.cuh) when the Cauchy search ends inside a group of equal
breakpoints.  Sample problem with l_odd = 2.5, x0 = 5 (iteration 2 ends inside a tie group), replay on and off."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import lbfgsb_b200

for n in (100_000, 1_000_000, 2_000_000):
    for limit in (1 << 21, 0):
        m = 5
        dev = torch.device("cuda")
        x = torch.full((n,), 5.0, dtype=torch.float64, device=dev)
        l = torch.full((n,), -100.0, dtype=torch.float64, device=dev); l[0::2] = 2.5
        u = torch.full((n,), 100.0, dtype=torch.float64, device=dev)
        nbd = torch.full((n,), 2, dtype=torch.int32, device=dev)
        g = torch.zeros_like(x)
        prob = lbfgsb_b200.DeviceProblem(n, m, np.float64)
        prob.set_tie_limit(limit)
        fg = lbfgsb_b200.RosenbrockDevice(np.float64)
        prob.profile(True)
        t_it = []
        t0 = time.perf_counter()
        while True:
            prob.setulb_dev(x, l, u, nbd, g, 0.0, 0.0)
            t = prob.task_str()
            if t[:2] == "FG":
                prob.f[0] = fg(x, g)
            elif t[:5] == "NEW_X":
                torch.cuda.synchronize()
                t_it.append((int(prob.isave[29]), int(prob.isave[32]), round((time.perf_counter() - t0) * 1e3, 2)))
                t0 = time.perf_counter()
                if prob.isave[29] >= 4:
                    break
            else:
                break
        pr = prob.profile_read()
        walk = {k: (round(v["ms"], 2), v["calls"]) for k, v in pr.items() if k.startswith("walk") and v["calls"]}
        print("n=%d tie_limit=%d: (iter, nseg, ms) %s  walk families (ms, calls) %s  tie_stats (replays, not replayed) %s" % (
            n, limit, t_it, walk, prob.tie_stats()), flush=True)
        prob.close()
