"""Development aid / regression fixture generator: bit-level digest of the engine's per-iterate trace.

The engine's reductions have a fixed shape (include/lbfgsb_b200_shape.h), so for a given problem every
iterate is bit-reproducible from run to run, from GPU to GPU and across kernel re-organisations that keep
the shape (fusing two passes into one must not change a single bit).  This tool runs a fixed list of
problems through the device-pointer entry point and prints, per problem, the iteration count and a
SHA-256 over the raw bits of (iter, nfgv, nseg, nfree, nact, nenter, nleave, iword, iback, col, nskip,
f, |proj g|, stp, theta, active-set hash) of every iterate.

  python tools/trace_digest.py [out.json]            # run on a GPU box
  tests/test_gpu_regression.py compares a fresh run with tests/golden/gpu_trace_digest.json
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# (name, n, m, dtype, l_odd, iteration budget, factr, pgtol)
CASES = [
    ("driver1_n25_m5_f64", 25, 5, "f64", 1.0, 100, 1e7, 1e-5),
    ("n1000_m10_l1.1_f64", 1000, 10, "f64", 1.1, 40, 0.0, 0.0),
    ("n100001_m5_l1.0_f64", 100001, 5, "f64", 1.0, 30, 0.0, 0.0),
    ("n1000000_m10_l1.1_f64", 1000000, 10, "f64", 1.1, 40, 0.0, 0.0),
    ("n300000_m20_l1.1_f64", 300000, 20, "f64", 1.1, 30, 0.0, 0.0),
    ("n4000_m20_l1.1_f32", 4000, 20, "f32", 1.1, 30, 0.0, 0.0),
    ("n1000000_m10_l1.1_f32", 1000000, 10, "f32", 1.1, 25, 0.0, 0.0),
    ("n200000_m3_unbounded_f64", 200000, 3, "f64", None, 25, 0.0, 0.0),
]


def run_case(case):
    import torch
    import lbfgsb_b200
    name, n, m, dt, l_odd, budget, factr, pgtol = case
    npdt = np.float64 if dt == "f64" else np.float32
    tdt = torch.float64 if dt == "f64" else torch.float32
    dev = torch.device("cuda", 0)
    x = torch.full((n,), 3.0, dtype=tdt, device=dev)
    l = torch.full((n,), -100.0, dtype=tdt, device=dev)
    u = torch.full((n,), 100.0, dtype=tdt, device=dev)
    if l_odd is None:
        nbd = torch.zeros(n, dtype=torch.int32, device=dev)
    else:
        l[0::2] = l_odd
        nbd = torch.full((n,), 2, dtype=torch.int32, device=dev)
    g = torch.zeros_like(x)
    prob = lbfgsb_b200.DeviceProblem(n, m, npdt)
    fg = lbfgsb_b200.RosenbrockDevice(npdt)
    h = hashlib.sha256()
    rows = 0
    fmt = "<11q4dQ"
    while True:
        prob.setulb_dev(x, l, u, nbd, g, factr, pgtol)
        t = prob.task_str()
        if t[:2] == "FG":
            if dt == "f64" and os.environ.get("LBFGSB_DIGEST_FUSED_FG") == "1":
                prob.f[0] = prob.fused_fg(0, x, g, l, u, nbd)     # objective kernel with the line-search epilogue
            else:
                prob.f[0] = fg(x, g)
        elif t[:5] == "NEW_X":
            hh, _ = prob.active_set_hash()
            i, d = prob.isave, prob.dsave
            h.update(struct.pack(fmt, int(i[29]), int(i[33]), int(i[32]), int(i[37]), int(i[38]), int(i[40]),
                                 int(n + 1 - i[39]), int(i[36]), int(i[24]), int(i[27]), int(i[25]),
                                 float(prob.f[0]), float(d[12]), float(d[13]), float(d[0]), hh))
            rows += 1
            if i[29] >= budget:
                break
        else:
            break
    xs = x.double()
    out = {"iterations": rows, "task": prob.task_str(), "sha256": h.hexdigest(),
           "f_final": float(prob.f[0]), "x_sum": float(xs.sum().item())}
    prob.close()
    return out


def main():
    res = {}
    for c in CASES:
        res[c[0]] = run_case(c)
        print(c[0], json.dumps(res[c[0]]), flush=True)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as fh:
            json.dump(res, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
