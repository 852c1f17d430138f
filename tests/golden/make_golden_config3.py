"""Records the oracle's first iterates of BASELINE.json configs[2] AT FULL SIZE (n = 1e8, m = 10, bounded extended
Rosenbrock with the odd lower bound at 1.1, factr = pgtol = 0) into tests/golden/config3_n1e8_trace.json.

The oracle (oracle/lbfgsb_oracle.cpp, 64-bit offsets) runs in device-order summation mode -- the mode the GPU is gated
against -- and, with --reference-order, in the reference's own order; the discrete trace of the two is identical
(checked here) and the drift of f between them is stored next to the trace.  About 25 GB of host memory, 14 min (device order) + 5 min (reference order) on one core; run once, the JSON is committed and tests/test_gpu_fullsize.py compares the GPU with it.

    python tests/golden/make_golden_config3.py [iterations=16] [--reference-order]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import harness as H  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

N, M, L_ODD = 100_000_000, 10, 1.1


def run(mode, iters):
    O.set_sum_mode(mode)
    try:
        x, l, u, nbd = H.rosenbrock_problem(N, l_odd=L_ODD)
        t0 = time.time()
        tr = H.run_driver(O.OracleSetulb(), O.rosenbrock_fg, N, M, x, l, u, nbd, 0.0, 0.0, stop=H.iteration_budget_stop(iters))
        print("mode %d: %d iterates in %.0f s" % (mode, len(tr[0]), time.time() - t0), flush=True)
        return tr[0]
    finally:
        O.set_sum_mode(0)


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16
    out_path = os.path.join(HERE, "config3_n1e8_trace.json")
    dev = run(1, iters)
    out = {"n": N, "m": M, "l_odd": L_ODD, "factr": 0.0, "pgtol": 0.0, "sum_mode": "device order (include/lbfgsb_b200_shape.h)",
           "fields": list(H.TRACE_FIELDS), "iterates": dev}
    if "--reference-order" in sys.argv:
        ref = run(0, iters)
        drift = []
        for a, b in zip(ref, dev):
            for k in ("iter", "nfgv", "nseg", "nact", "nfree", "nenter", "nleave", "iword", "iback", "col", "nskip", "hash", "hcount"):
                assert a[k] == b[k], (k, a, b)
            drift.append(abs(a["f"] - b["f"]) / max(abs(a["f"]), 1e-300))
        out["reference_order"] = {"discrete_trace_identical": True, "rel_drift_of_f_per_iterate": drift,
                                  "f": [a["f"] for a in ref], "sbgnrm": [a["sbgnrm"] for a in ref]}
    with open(out_path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
