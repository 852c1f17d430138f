"""Equal breakpoints at the exit of the generalized-Cauchy-point search, on the CPU oracle alone.

hpsolb (src/lbfgsb.f90:2079-2157) pops equal keys in the order of its heap; when the search ends inside such a group
(:1416 with dt = 0), the members popped so far are fixed at their bounds and the others stay free.  The oracle's default
mode is the reference's heap; `set_tie_mode(1)` takes ties in variable order (what a stable sort gives, and what the
CUDA engine does when its heap replay is switched off or out of range).  These tests pin what the GPU parity tests of
tests/test_gpu_rare_paths.py rely on: the two orders give the same point and the same counts but a different active
set; the same problems also exercise subsm's backtrack (:2830-2879)."""
import numpy as np
import pytest

import harness as H
from oracle import oracle_py as O


def _run(n, m, l_odd, x0, tie_mode, budget):
    O.set_tie_mode(tie_mode)
    O.event_counts()
    try:
        x, l, u, nbd = H.rosenbrock_problem(n, l_odd=l_odd, x0=x0)
        r = H.run_driver(O.OracleSetulb(), O.rosenbrock_fg, n, m, x, l, u, nbd, 0.0, 0.0,
                         stop=H.iteration_budget_stop(budget))
    finally:
        O.set_tie_mode(0)
    return r, O.event_counts()


@pytest.mark.parametrize("n,m,l_odd,x0", [(1000, 5, 2.0, 3.0), (3001, 3, 2.5, 5.0)])
def test_tie_order_changes_only_the_active_set_at_the_exit(n, m, l_odd, x0):
    heap, ev_heap = _run(n, m, l_odd, x0, 0, 40)
    var, ev_var = _run(n, m, l_odd, x0, 1, 40)
    a, b = heap[0][1], var[0][1]          # iterate 2: its Cauchy search ended inside a tie group
    assert a["nseg"] == b["nseg"] > 2 and a["nact"] == b["nact"] and a["nfgv"] == b["nfgv"]
    assert a["f"] == b["f"]               # the same point: tied variables sit on their bounds either way
    assert a["hash"] != b["hash"]         # but different members of the group are marked active
    assert heap[0][0]["hash"] == var[0][0]["hash"]   # iterate 1 (every breakpoint passed) does not depend on the order
    assert ev_heap[0] >= 1                # subsm's backtrack (:2830) runs somewhere in the 40 iterations


def test_tie_mode_does_not_touch_problems_without_such_an_exit():
    heap, _ = _run(1000, 10, 1.1, 3.0, 0, 12)
    var, _ = _run(1000, 10, 1.1, 3.0, 1, 12)
    for a, b in zip(heap[0], var[0]):
        assert a == b
