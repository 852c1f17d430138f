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


def test_fortran_module_binds_only_exported_symbols():
    """fortran/lbfgsb_b200_module.F90 cannot be compiled in this image (no Fortran compiler): at least every
    bind(C, name='...') of its interface block must name a symbol that the library exports and the header declares."""
    import lbfgsb_b200
    src = open(os.path.join(ROOT, "fortran", "lbfgsb_b200_module.F90")).read()
    hdr = open(os.path.join(ROOT, "include", "lbfgsb_b200.h")).read()
    names = sorted(set(re.findall(r"bind\(C,\s*name='([A-Za-z0-9_]+)'\)", src)))
    assert len(names) >= 12
    L = lbfgsb_b200.lib()
    for nm in names:
        assert hasattr(L, nm), nm
        assert re.search(r"\b%s\s*\(" % nm, hdr), nm


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


def test_host_heap_replay_takes_ties_in_the_reference_order():
    """The engine's host-thread heap replay (Engine::heap_group_order; used for long breakpoint lists and, on every rank,
    for sharded workspaces) against the oracle's verbatim hpsolb driven as cauchy drives it (src/lbfgsb.f90:1384-1401,
    2079-2157): the members of a group of equal breakpoints come out in the same order.  Host code only: runs without a GPU."""
    import ctypes as C
    import lbfgsb_b200
    from oracle import oracle_py as O
    L = lbfgsb_b200.lib()
    L.lbfgsb_test_host_heap_group_f64.argtypes = [C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    for nb, distinct in ((1, 1), (2, 1), (9, 2), (100, 3), (1000, 1), (4097, 7), (60000, 11), (60000, 60000)):
        vals = np.sort(rng.uniform(0.1, 5.0, distinct))
        t = vals[rng.integers(0, distinct, nb)].astype(np.float64)
        for tk in set([float(t.min()), float(np.median(t)), float(t.max())]):
            out = np.empty(nb, dtype=np.int32)
            cnt = C.c_int64(0)
            assert L.lbfgsb_test_host_heap_group_f64(nb, t.ctypes.data_as(C.c_void_p), tk, out.ctypes.data_as(C.c_void_p), C.byref(cnt)) == 0
            got = list(out[:cnt.value])
            # the reference: first minimum (lowest index among ties) taken before the heap exists, replaced by the last
            # entry; then hpsolb builds the heap on its first call and pops the least member into t(nleft) on every call
            tt = t.copy(); io = np.arange(nb, dtype=np.int32)
            ibp = int(np.argmin(tt))
            ref = [ibp] if tt[ibp] == tk else []
            if ibp != nb - 1:
                tt[ibp] = tt[nb - 1]; io[ibp] = io[nb - 1]
            nleft, k = nb - 1, 0
            while nleft > 0:
                O.lib().oracle_hpsolb_f64(C.c_int64(nleft), tt.ctypes.data_as(C.c_void_p), io.ctypes.data_as(C.c_void_p), C.c_int64(0 if k == 0 else 1))
                v, var = tt[nleft - 1], int(io[nleft - 1])
                nleft -= 1; k += 1
                if v > tk:
                    break
                if v == tk:
                    ref.append(var)
            assert got == ref, (nb, distinct, tk, got[:10], ref[:10])
