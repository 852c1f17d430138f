"""Protocol edges of the reverse-communication interface on the CUDA engine (C-ABI host twin and
device-pointer variant): user STOP with and without the 'CPU' restore (src/lbfgsb.f90:565-571,
test/driver3.f90:152-182), reading the previous iterate t, restart on an unknown task (:573-575),
misuse errors, and BASELINE.json config 2 (n = 1e6, m = 5) against the oracle."""
import numpy as np
import pytest

import harness as H
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


def _state(n, dtype=np.float64):
    return dict(task=H.make_task("START"), csave=H.make_task(""), lsave=np.zeros(4, np.int32),
                isave=np.zeros(44, np.int32), dsave=np.zeros(29, dtype), f=np.zeros(1, dtype), g=np.zeros(n, dtype))


def _drive(setulb, n, m, x, l, u, nbd, st, until_iter, factr=0.0, pgtol=0.0):
    wa, iwa = setulb.workspace(n, m)
    xs = {}
    while True:
        setulb(n, m, x, l, u, nbd, st["f"], st["g"], factr, pgtol, wa, iwa, st["task"], -1, st["csave"], st["lsave"],
               st["isave"], st["dsave"])
        ts = H.task_str(st["task"])
        if ts[:2] == "FG":
            st["f"][0] = O.rosenbrock_fg(x, st["g"])
        elif ts[:5] == "NEW_X":
            xs[int(st["isave"][29])] = (x.copy(), st["g"].copy(), float(st["f"][0]))
            if st["isave"][29] >= until_iter:
                return xs, wa, iwa
        else:
            return xs, wa, iwa


@pytest.mark.parametrize("impl", ["gpu", "oracle"])
def test_stop_cpu_restores_previous_iterate(impl):
    """task = 'STOP: CPU ...' at NEW_X puts the previous iterate (t, r, fold) back into x, g, f."""
    import lbfgsb_b200
    n, m = 1000, 10
    x, l, u, nbd = H.rosenbrock_problem(n)
    s = lbfgsb_b200.HostSetulb() if impl == "gpu" else O.OracleSetulb()
    st = _state(n)
    xs, wa, iwa = _drive(s, n, m, x, l, u, nbd, st, 7)
    assert H.task_str(st["task"]) == "NEW_X"
    if impl == "gpu":
        t = s.previous_x(n)                      # what driver3 reads out of wa(lt:lt+n-1)
        assert np.array_equal(t, xs[6][0])
    st["task"][:] = H.make_task("STOP: CPU EXCEEDING THE TIME LIMIT")
    s(n, m, x, l, u, nbd, st["f"], st["g"], 0.0, 0.0, wa, iwa, st["task"], -1, st["csave"], st["lsave"], st["isave"], st["dsave"])
    assert H.task_str(st["task"]).startswith("STOP: CPU")
    assert np.array_equal(x, xs[6][0]) and np.array_equal(st["g"], xs[6][1]) and st["f"][0] == xs[6][2]
    s.release(st["isave"])


def test_plain_stop_leaves_everything_and_frees():
    import lbfgsb_b200
    n, m = 500, 5
    x, l, u, nbd = H.rosenbrock_problem(n)
    s = lbfgsb_b200.HostSetulb()
    st = _state(n)
    xs, wa, iwa = _drive(s, n, m, x, l, u, nbd, st, 4)
    xk = x.copy()
    st["task"][:] = H.make_task("STOP: THE USER IS DONE")
    s(n, m, x, l, u, nbd, st["f"], st["g"], 0.0, 0.0, wa, iwa, st["task"], -1, st["csave"], st["lsave"], st["isave"], st["dsave"])
    assert H.task_str(st["task"]) == "STOP: THE USER IS DONE" and np.array_equal(x, xk)
    assert s.engine() is None        # the workspace was released


def test_unknown_task_restarts_with_fg_start():
    """Any other task on re-entry is answered with 'FG_START' (start(), :573-575, :884-890)."""
    import lbfgsb_b200
    n, m = 200, 4
    x, l, u, nbd = H.rosenbrock_problem(n)
    s = lbfgsb_b200.HostSetulb()
    st = _state(n)
    _, wa, iwa = _drive(s, n, m, x, l, u, nbd, st, 3)
    st["task"][:] = H.make_task("WHATEVER")
    s(n, m, x, l, u, nbd, st["f"], st["g"], 0.0, 0.0, wa, iwa, st["task"], -1, st["csave"], st["lsave"], st["isave"], st["dsave"])
    assert H.task_str(st["task"]) == "FG_START"
    s.release(st["isave"])


def test_reentry_without_start_is_an_error():
    import lbfgsb_b200
    n, m = 50, 3
    x, l, u, nbd = H.rosenbrock_problem(n)
    st = _state(n)
    st["task"][:] = H.make_task("FG_LNSRCH")
    with pytest.raises(lbfgsb_b200.LbfgsbB200Error):
        lbfgsb_b200.setulb(n, m, x, l, u, nbd, st["f"], st["g"], 1e7, 1e-5, None, None, st["task"], -1, st["csave"],
                           st["lsave"], st["isave"], st["dsave"])


def test_unaligned_device_pointer_is_refused():
    import torch
    import lbfgsb_b200
    n, m = 1024, 5
    buf = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
    x = buf[1:]                                  # 8-byte aligned only
    l = torch.zeros(n, dtype=torch.float64, device="cuda")
    u = torch.ones(n, dtype=torch.float64, device="cuda")
    nbd = torch.full((n,), 2, dtype=torch.int32, device="cuda")
    g = torch.zeros(n, dtype=torch.float64, device="cuda")
    prob = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    with pytest.raises(lbfgsb_b200.LbfgsbB200Error):
        prob.setulb_dev(x, l, u, nbd, g, 1e7, 1e-5)
    prob.close()


def test_config2_n1e6_m5_full_trace_vs_oracle():
    """BASELINE.json config 2: bounded extended Rosenbrock n=1e6, m=5, factr=1e7, pgtol=1e-5 -- same iteration
    count to convergence, same per-iterate free/active classification, same final active set."""
    import lbfgsb_b200
    n, m = 1_000_000, 5
    x, l, u, nbd = H.rosenbrock_problem(n)
    gpu = H.run_driver(lbfgsb_b200.HostSetulb(), O.rosenbrock_fg, n, m, x, l, u, nbd, 1.0e7, 1.0e-5)
    x2, l, u, nbd = H.rosenbrock_problem(n)
    O.set_sum_mode(1)
    try:
        ref = H.run_driver(O.OracleSetulb(), O.rosenbrock_fg, n, m, x2, l, u, nbd, 1.0e7, 1.0e-5)
    finally:
        O.set_sum_mode(0)
    assert gpu[1] == ref[1] and gpu[1].startswith("CONVERGENCE")
    assert len(gpu[0]) == len(ref[0]) == 31
    assert max(r["nseg"] for r in gpu[0]) == n          # iteration 1 walks every breakpoint
    for a, b in zip(gpu[0], ref[0]):
        for k in ("iter", "nfgv", "nseg", "nact", "nfree", "nenter", "nleave", "iword", "iback", "col", "hash", "hcount"):
            assert a[k] == b[k], (k, a, b)
        tol = 1e-10 if b["iter"] <= 10 else 1e-6
        assert abs(a["f"] - b["f"]) <= tol * abs(b["f"]), (a, b)
        assert abs(a["sbgnrm"] - b["sbgnrm"]) <= tol * abs(b["sbgnrm"]), (a, b)
    assert np.max(np.abs(gpu[2] - ref[2])) <= 1e-6
