"""Development aid: iteration rate of BASELINE.json configs[4] (driver3-style bounded problem, n = 4e8, m = 20, REAL32)
on one B200 -- the widest S/Y history (40 columns of 1.6 GB).  Prints ms per iteration over the iterations whose
history is full (col = m) and the per-kernel-family times of those iterations."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import lbfgsb_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
m = 20
warm, steps = m + 4, 10
dev = torch.device("cuda")
x = torch.full((n,), 3.0, dtype=torch.float32, device=dev)
l = torch.full((n,), -100.0, dtype=torch.float32, device=dev); l[0::2] = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
u = torch.full((n,), 100.0, dtype=torch.float32, device=dev)
nbd = torch.full((n,), 2, dtype=torch.int32, device=dev)
g = torch.zeros_like(x)
st = torch.cuda.current_stream().cuda_stream
prob = lbfgsb_b200.DeviceProblem(n, m, np.float32, stream=st)
fg = lbfgsb_b200.RosenbrockDevice(np.float32, stream=st)
import time
prob.profile(True)
rows = []
pr0 = None
nfg = 0
torch.cuda.synchronize(); t0 = time.perf_counter(); nfg_last = 0
while True:
    prob.setulb_dev(x, l, u, nbd, g, 0.0, 0.0)
    t = prob.task_str()
    if t[:2] == "FG":
        prob.f[0] = fg(x, g); nfg += 1
    elif t[:5] == "NEW_X":
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        rows.append((int(prob.isave[29]), int(prob.isave[27]), int(prob.isave[32]), int(prob.isave[37]), nfg - nfg_last, (t1 - t0) * 1e3, float(prob.f[0])))
        t0 = t1; nfg_last = nfg
        if int(prob.isave[29]) == warm and pr0 is None:
            pr0 = prob.profile_read()
        if int(prob.isave[29]) >= 40:
            break
    else:
        print("ended:", t)
        break
w = 4
print("config 5: n=%d m=%d REAL32 -- iter, col, nseg, nfree, f/g evals, ms (wall, whole iteration incl. the f/g kernel), f" % (n, m))
for r in rows:
    it, col, nseg, nfree, nf, ms, f = r
    canon = n * (35 * w + 40) + nfree * (12 * w + 20) + 2 * col * w * (2 * n + nfree + 3 * nfree)
    print("  %3d  col %2d  nseg %9d  nfree %9d  fg %d  %8.2f ms  canonical %6.1f GB -> %6.0f GB/s   f %.6e" % (it, col, nseg, nfree, nf, ms, canon / 1e9, canon / ms / 1e6, f))
pr = prob.profile_read()
print("kernel families over the whole run (total ms, calls):", {k: (round(v["ms"], 1), v["calls"]) for k, v in pr.items() if v["ms"] > 5})
if pr0 is not None:
    print("kernel families after iteration %d, history full (ms per call, calls):" % warm,
          {k: (round((v["ms"] - pr0.get(k, {"ms": 0})["ms"]) / max(1, v["calls"] - pr0.get(k, {"calls": 0})["calls"]), 2),
               v["calls"] - pr0.get(k, {"calls": 0})["calls"]) for k, v in pr.items() if v["ms"] - pr0.get(k, {"ms": 0})["ms"] > 1})
