"""the library-owned task loop (lbfgsb_minimize_dev_*, the reference's
@todo src/lbfgsb.f90:36-37, loop shape test/driver1.f90:263-292 / driver2.f90:174-181) and checkpoint / resume of
the device workspace (the reference's state is the caller's wa/iwa, test/driver3.f90:152-182)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(n, l_odd=1.1):
    import torch
    x = torch.full((n,), 3.0, dtype=torch.float64, device="cuda")
    l = torch.full((n,), -100.0, dtype=torch.float64, device="cuda")
    l[0::2] = l_odd
    u = torch.full((n,), 100.0, dtype=torch.float64, device="cuda")
    nbd = torch.full((n,), 2, dtype=torch.int32, device="cuda")
    return x, l, u, nbd, torch.zeros_like(x)


def _row(prob):
    return (int(prob.isave[29]), int(prob.isave[33]), int(prob.isave[32]), int(prob.isave[37]), float(prob.f[0]).hex(),
            float(prob.dsave[12]).hex(), float(prob.dsave[0]).hex(), prob.active_set_hash()[0])


def _loop(prob, fg, x, l, u, nbd, g, factr, pgtol, until_iter):
    rows = []
    while True:
        prob.setulb_dev(x, l, u, nbd, g, factr, pgtol)
        t = prob.task_str()
        if t[:2] == "FG":
            prob.f[0] = fg(x, g)
        elif t[:5] == "NEW_X":
            rows.append(_row(prob))
            if prob.isave[29] >= until_iter:
                break
        else:
            break
    return rows


def test_minimize_owns_the_loop_and_matches_the_caller_loop():
    import lbfgsb_b200
    n, m = 30000, 7
    fg = lbfgsb_b200.RosenbrockDevice(np.float64)
    x, l, u, nbd, g = _problem(n)
    a = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    _loop(a, fg, x, l, u, nbd, g, 1e7, 1e-5, 10 ** 9)
    x2, l, u, nbd, g2 = _problem(n)
    b = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    rc = b.minimize(x2, l, u, nbd, g2, fg, 1e7, 1e-5)
    assert rc == 0 and b.task_str() == a.task_str() and b.task_str().startswith("CONVERGENCE")
    assert float(a.f[0]).hex() == float(b.f[0]).hex()
    assert list(a.isave[21:44]) == list(b.isave[21:44])
    import torch
    assert torch.equal(x, x2)
    a.close(); b.close()


def test_minimize_stop_limits():
    import lbfgsb_b200
    n, m = 5000, 5
    fg = lbfgsb_b200.RosenbrockDevice(np.float64)
    x, l, u, nbd, g = _problem(n)
    p = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    assert p.minimize(x, l, u, nbd, g, fg, 0.0, 0.0, max_iter=7) == 0
    assert p.task_str().startswith("STOP: TOTAL NO. of ITERATIONS") and int(p.isave[29]) == 7
    x, l, u, nbd, g = _problem(n)
    q = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    assert q.minimize(x, l, u, nbd, g, fg, 0.0, 0.0, max_fg=9) == 0
    assert q.task_str().startswith("STOP: TOTAL NO. of f AND g") and int(q.isave[33]) >= 9

    def bad(xx, gg):
        raise RuntimeError("objective failed")
    x, l, u, nbd, g = _problem(n)
    r = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    assert r.minimize(x, l, u, nbd, g, bad, 0.0, 0.0) == 2 and r.task_str().startswith("STOP: THE OBJECTIVE CALLBACK FAILED")
    for o in (p, q, r):
        o.close()


def test_checkpoint_resume_continues_bit_identically(tmp_path):
    import lbfgsb_b200
    n, m = 40001, 10
    fg = lbfgsb_b200.RosenbrockDevice(np.float64)
    x, l, u, nbd, g = _problem(n)
    a = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    _loop(a, fg, x, l, u, nbd, g, 0.0, 0.0, 9)
    path = os.path.join(str(tmp_path), "ws.ckp")
    a.checkpoint_write(path)
    saved = dict(x=x.clone(), g=g.clone(), f=a.f.copy(), task=a.task.copy(), csave=a.csave.copy(), lsave=a.lsave.copy(),
                 isave=a.isave.copy(), dsave=a.dsave.copy())
    rest_a = _loop(a, fg, x, l, u, nbd, g, 0.0, 0.0, 20)
    a.close()
    assert os.path.getsize(path) > (2 * m + 5) * n * 8

    b = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    b.checkpoint_read(path)
    xb, gb = saved["x"].clone(), saved["g"].clone()
    b.f[:] = saved["f"]; b.task[:] = saved["task"]; b.csave[:] = saved["csave"]; b.lsave[:] = saved["lsave"]
    b.isave[:] = saved["isave"]; b.dsave[:] = saved["dsave"]
    rest_b = _loop(b, fg, xb, l, u, nbd, gb, 0.0, 0.0, 20)
    assert len(rest_a) == 11 and rest_a == rest_b
    import torch
    assert torch.equal(x, xb)
    # the previous iterate `t` (what driver3 reads out of wa, test/driver3.f90:173-175) survives the round trip
    assert b.vector(3).shape[0] == n
    b.close()

    c = lbfgsb_b200.DeviceProblem(n + 1, m, np.float64)
    with pytest.raises(lbfgsb_b200.LbfgsbB200Error):
        c.checkpoint_read(path)
    c.close()


@pytest.mark.parametrize("n,m,factr,pgtol,max_iter", [(30000, 7, 1e7, 1e-5, 0), (1000, 10, 0.0, 0.0, 40), (200001, 5, 0.0, 0.0, 25)])
def test_minimize_graph_keeps_the_loop_on_the_device_and_matches_the_caller_loop(n, m, factr, pgtol, max_iter):
    """lbfgsb_minimize_graph_dev_f64: one CUDA-graph launch per iteration step (objective + FG_LNSRCH entry + NEW_X entry),
    everything off the common path handed to the general pipeline -- final x, f, isave(22:44) and the task equal to the
    caller-driven loop bit for bit (the first iterations of these problems walk breakpoints and move variables in and out of
    the free set, later ones run as graph launches only)."""
    import torch
    import lbfgsb_b200
    fg = lbfgsb_b200.RosenbrockDevice(np.float64)
    x, l, u, nbd, g = _problem(n)
    a = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    if max_iter:
        rc = a.minimize(x, l, u, nbd, g, fg, factr, pgtol, max_iter=max_iter)
    else:
        _loop(a, fg, x, l, u, nbd, g, factr, pgtol, 10 ** 9)
    x2, l, u, nbd, g2 = _problem(n)
    b = lbfgsb_b200.DeviceProblem(n, m, np.float64)
    fg2 = lbfgsb_b200.RosenbrockDevice(np.float64)
    fg2._n = n
    halo = torch.zeros(2, dtype=torch.float64, device="cuda")
    rc = b.minimize_graph(x2, l, u, nbd, g2, fg2.enqueue(halo), factr, pgtol, max_iter=max_iter)
    assert rc == 0 and b.task_str() == a.task_str()
    assert float(a.f[0]).hex() == float(b.f[0]).hex()
    assert list(a.isave[21:44]) == list(b.isave[21:44])
    assert torch.equal(x, x2)
    steps, per = b.graph_stats()
    assert steps >= int(b.isave[29]) // 2 and per > 0, (steps, per, int(b.isave[29]))
    a.close(); b.close()
