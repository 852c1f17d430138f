"""Development aid: rate of the batched small-problem path (include/lbfgsb_b200.h section 6) on driver1-sized problems
(n = 25, m = 5, factr = 1e7, pgtol = 1e-5, perturbed starting points): problems per second to convergence, next to the CPU
oracle solving the same problems one after the other.   python tools/batch_rate.py [nprob=1000] [n=25] [m=5]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def problems(nprob, n, seed=7):
    rng = np.random.default_rng(seed)
    x0 = 3.0 + 0.5 * rng.uniform(-1.0, 1.0, (nprob, n))
    l = np.empty((nprob, n)); u = np.full((nprob, n), 100.0)
    l[:, 0::2] = 1.0; l[:, 1::2] = -100.0
    nbd = np.full((nprob, n), 2, dtype=np.int32)
    return x0, l, u, nbd


def gpu_rate(nprob, n, m, factr=1.0e7, pgtol=1.0e-5, repeat=3):
    import torch
    import lbfgsb_b200
    x0, l, u, nbd = problems(nprob, n)
    ld, ud, nd = (torch.from_numpy(a).cuda() for a in (l, u, nbd))
    best = None
    for _ in range(repeat):
        xd = torch.from_numpy(x0.copy()).cuda()
        gd = torch.zeros_like(xd); fd = torch.zeros(nprob, dtype=torch.float64, device="cuda")
        b = lbfgsb_b200.BatchProblem(nprob, n, m, np.float64)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        calls = b.solve(xd, ld, ud, nd, fd, gd, factr, pgtol)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        its = int(b.isave[:, 29].sum()); nfg = int(b.isave[:, 33].sum())
        conv = int(sum(1 for p in range(nprob) if b.task_str(p).startswith("CONV")))
        b.close()
        if best is None or dt < best["seconds"]:
            best = {"nprob": nprob, "n": n, "m": m, "seconds": dt, "problems_per_s": nprob / dt, "iterations_per_s": its / dt,
                    "calls": calls, "iterations_total": its, "fg_total": nfg, "converged": conv,
                    "us_per_call": dt / calls * 1e6}
    return best


def cpu_rate(nprob, n, m, factr=1.0e7, pgtol=1.0e-5):
    import harness as H
    from oracle import oracle_py as O
    x0, l, u, nbd = problems(nprob, n)
    t0 = time.perf_counter()
    its = 0
    for p in range(nprob):
        tr = H.run_driver(O.OracleSetulb(), O.rosenbrock_fg, n, m, x0[p].copy(), l[p], u[p], nbd[p], factr, pgtol, want_hash=False)
        its += len(tr[0])
    dt = time.perf_counter() - t0
    return {"nprob": nprob, "seconds": dt, "problems_per_s": nprob / dt, "iterations_per_s": its / dt, "cores": 1,
            "note": "CPU oracle port through its Python driver loop, one problem after the other"}


if __name__ == "__main__":
    nprob = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    out = {"gpu": gpu_rate(nprob, n, m), "cpu": cpu_rate(min(nprob, 1000), n, m)}
    print(json.dumps(out))
