"""Development aid: run one problem through the CUDA engine and the CPU oracle (device-order sums)
and print the two per-iterate traces side by side.  Not used by the product.

  python tools/trace_compare.py N M [l_odd] [max_iter] [factr] [pgtol]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import harness as H  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def main():
    import lbfgsb_b200
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    l_odd = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    mi = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    factr = float(sys.argv[5]) if len(sys.argv) > 5 else 1e7
    pgtol = float(sys.argv[6]) if len(sys.argv) > 6 else 1e-5
    stop = H.iteration_budget_stop(mi)
    x, l, u, nbd = H.rosenbrock_problem(n, l_odd=l_odd)
    O.set_sum_mode(1)
    ref = H.run_driver(O.OracleSetulb(), O.rosenbrock_fg, n, m, x, l, u, nbd, factr, pgtol, stop=stop)
    O.set_sum_mode(0)
    x, l, u, nbd = H.rosenbrock_problem(n, l_odd=l_odd)
    try:
        gpu = H.run_driver(lbfgsb_b200.HostSetulb(), O.rosenbrock_fg, n, m, x, l, u, nbd, factr, pgtol, stop=stop,
                           max_calls=4 * mi + 50)
    except Exception as e:  # noqa: BLE001
        print("GPU run failed:", e)
        gpu = ([], "FAILED", None, 0.0, None, None)
    keys = ("iter", "nfgv", "nseg", "nact", "nfree", "nenter", "nleave", "iword", "iback", "col", "nskip")
    print("task gpu: %r   oracle: %r" % (gpu[1], ref[1]))
    for i in range(max(len(gpu[0]), len(ref[0]))):
        a = gpu[0][i] if i < len(gpu[0]) else None
        b = ref[0][i] if i < len(ref[0]) else None
        for tag, r in (("gpu", a), ("ref", b)):
            if r is None:
                print("%s  --" % tag)
                continue
            print("%s %s f=%.15e pg=%.15e stp=%.6e th=%.6e h=%016x" % (
                tag, " ".join("%s=%d" % (k, r[k]) for k in keys), r["f"], r["sbgnrm"], r["stp"], r["theta"],
                r["hash"] or 0))
        if a and b:
            bad = [k for k in keys + ("hash",) if a[k] != b[k]]
            rel = abs(a["f"] - b["f"]) / max(abs(b["f"]), 1e-300)
            print("    rel df=%.2e %s" % (rel, ("MISMATCH " + ",".join(bad)) if bad else ""))


if __name__ == "__main__":
    main()
