"""CPU tests of the N > 1 host logic (gloo, world_size 2 and 3) and of the C-ABI symbol table."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_halo_and_partition_under_gloo(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "gloo_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "GLOO_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_shard_numpy_matches_oracle_fg():
    from lbfgsb_b200 import sharded
    from oracle import oracle_py as O
    rng = np.random.default_rng(3)
    x = rng.uniform(-2, 2, 501)
    g = np.zeros_like(x)
    f = O.rosenbrock_fg(x, g)
    f2, g2 = sharded.rosenbrock_shard_numpy(x, True, True, 0.0, 0.0)
    assert abs(f - f2) <= 1e-12 * abs(f)
    assert np.allclose(g, g2, rtol=1e-13, atol=0)


def test_quadratic_problem_oracle_half_active():
    """BASELINE.json configs[3] at a size the oracle solves in a second: the quadratic is convex, the oracle
    converges, about half of the variables end on the lower bound, and the result is the box-constrained
    minimiser (projected gradient zero to pgtol)."""
    import harness as H
    from lbfgsb_b200 import sharded
    from oracle import oracle_py as O
    n, m = 4000, 10

    def fg(x, g):
        f, gg = sharded.quadratic_shard_numpy(x, 0, 0, 0.0, 0.0)
        g[:] = gg
        return f
    x, l, u, nbd = sharded.quadratic_problem(n)
    tr, task, x, f, isave, dsave = H.run_driver(O.OracleSetulb(), fg, n, m, x, l, u, nbd, 1.0e1, 1.0e-8)
    assert task.startswith("CONVERGENCE"), task
    assert 0.4 * n < tr[-1]["nact"] < 0.6 * n
    fs = [r["f"] for r in tr]
    assert all(b <= a for a, b in zip(fs, fs[1:]))
    _, g = sharded.quadratic_shard_numpy(x, 0, 0, 0.0, 0.0)
    pg = np.where(g < 0, np.maximum(x - u, g), np.minimum(x - l, g))
    assert np.abs(pg).max() <= 1e-6
    # coefficients are a pure function of the global index: a shard sees the same numbers
    d0, b0 = sharded.quadratic_coefficients(0, n, 0)
    d1, b1 = sharded.quadratic_coefficients(1234, 2000, 0)
    assert np.array_equal(d0[1234:2000], d1) and np.array_equal(b0[1234:2000], b1)
    assert 2.1 <= d0.min() and d0.max() < 3.1 and -1.0 <= b0.min() and b0.max() < 1.0


def test_library_exports_every_declared_symbol():
    """No compute calls: only that the in-tree .so loads and exports include/lbfgsb_b200.h."""
    import lbfgsb_b200
    if not os.path.exists(lbfgsb_b200.SO_PATH):
        lbfgsb_b200.build()
    L = ctypes.CDLL(lbfgsb_b200.SO_PATH)
    hdr = open(os.path.join(ROOT, "include", "lbfgsb_b200.h")).read()
    names = sorted(set(re.findall(r"\b(lbfgsb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    for nm in names:
        assert hasattr(L, nm), nm
    L.lbfgsb_b200_version.restype = ctypes.c_int
    assert L.lbfgsb_b200_version() >= 100


def test_no_cpu_path_without_a_gpu():
    """Without a CUDA device the host twin must refuse (task = 'ERROR: NO CUDA DEVICE ...'), never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import harness as H
    import lbfgsb_b200
    n, m = 10, 3
    x, l, u, nbd = H.rosenbrock_problem(n)
    task = H.make_task("START")
    csave = H.make_task("")
    with pytest.raises(lbfgsb_b200.LbfgsbB200Error):
        lbfgsb_b200.setulb(n, m, x, l, u, nbd, np.zeros(1), np.zeros(n), 1e7, 1e-5, None, None, task, -1, csave,
                           np.zeros(4, np.int32), np.zeros(44, np.int32), np.zeros(29))
    assert H.task_str(task).startswith("ERROR: NO CUDA DEVICE")
    assert lbfgsb_b200.lib().lbfgsb_dev_create(10, 3, 8, None) is None


def test_product_does_not_import_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "lbfgsb_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liblbfgsb_oracle" not in src, fn
