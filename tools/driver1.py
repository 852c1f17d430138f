"""The reference's sample program test/driver1.f90 (n = 25, m = 5, iprint = 1, factr = 1e7, pgtol = 1e-5)
through the C-ABI host twin: the library prints what the Fortran library prints (host_print.h) and
writes the summary file.  Usage: python tools/driver1.py [iprint] [iteration_file] [n] [m]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import harness as H  # noqa: E402
import lbfgsb_b200  # noqa: E402


def fg(x, g):
    """test/driver1.f90:274-289"""
    n = x.shape[0]
    f = 0.25 * (x[0] - 1.0) ** 2
    f += float(((x[1:] - x[:-1] ** 2) ** 2).sum())
    f *= 4.0
    t1 = x[1] - x[0] ** 2
    g[0] = 2.0 * (x[0] - 1.0) - 16.0 * x[0] * t1
    for i in range(1, n - 1):
        t2 = t1
        t1 = x[i + 1] - x[i] ** 2
        g[i] = 8.0 * t2 - 16.0 * x[i] * t1
    g[n - 1] = 8.0 * t1
    return f


def main():
    iprint = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    itfile = sys.argv[2] if len(sys.argv) > 2 else "iterate.dat"
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    m = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    print("\n     Solving sample problem.\n      (f = 0.0 at the optimal solution.)\n", flush=True)
    x, l, u, nbd = H.rosenbrock_problem(n)
    from oracle import oracle_py as O   # the same f/g routine the parity tests feed to both sides
    H.run_driver(lbfgsb_b200.HostSetulb(iteration_file=itfile), O.rosenbrock_fg, n, m, x, l, u, nbd, 1.0e7, 1.0e-5,
                 iprint=iprint, want_hash=False)


if __name__ == "__main__":
    main()
