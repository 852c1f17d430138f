"""Reverse-communication driver shared by the tests and bench.py.

`run_driver` plays the role of the reference's sample programs
(/root/reference/test/driver1.f90:259-292, driver2.f90:112-190, driver3.f90:122-236):
it owns the `task` loop, evaluates f/g when task(1:2)=='FG', and records one trace
row per 'NEW_X' from isave/dsave exactly as driver2/driver3 print them.

`setulb` is any callable with the host-twin convention -- the oracle
(oracle/oracle_py.py: OracleSetulb) or the CUDA engine
(lbfgsb_b200.HostSetulb) -- so the same loop drives both sides of a parity test.
"""
import numpy as np


def make_task(s="START"):
    t = np.full(60, ord(" "), dtype=np.uint8)
    b = s.encode()
    t[: len(b)] = np.frombuffer(b, dtype=np.uint8)
    return t


def task_str(task):
    return bytes(task).decode().rstrip()


def rosenbrock_problem(n, dtype=np.float64, l_odd=1.0, x0=3.0):
    """Bounds and start of test/driver1.f90:233-251.  Odd (1-based) variables: [l_odd, 100];
    even: [-100, 100]; nbd = 2 everywhere; x0 = 3.  `l_odd=1.1` is BASELINE.json config 3
    (cuts off the unconstrained minimiser so ~50% of variables end at a bound)."""
    l = np.empty(n, dtype=dtype)
    u = np.full(n, 100.0, dtype=dtype)
    l[0::2] = l_odd
    l[1::2] = -100.0
    nbd = np.full(n, 2, dtype=np.int32)
    x = np.full(n, x0, dtype=dtype)
    return x, l, u, nbd


TRACE_FIELDS = ("iter", "nfgv", "nseg", "nact", "nfree", "nenter", "nleave", "iword", "iback",
                "col", "nskip", "nintol", "stp", "xstep", "theta", "sbgnrm", "f", "hash", "hcount")


def run_driver(setulb, fg, n, m, x, l, u, nbd, factr, pgtol, iprint=-1, stop=None, max_calls=100000,
               want_hash=True, keep_x=False):
    """Returns (trace rows, final task string, x, f, isave, dsave).

    stop(isave, dsave, f) -> str or None is the driver2/driver3-style user stop rule,
    evaluated at every NEW_X; a returned string is written into task (task(1:4)=='STOP').
    """
    dtype = x.dtype
    g = np.zeros(n, dtype=dtype)
    f = np.zeros(1, dtype=dtype)
    wa, iwa = setulb.workspace(n, m)
    task = make_task("START")
    csave = make_task("")
    lsave = np.zeros(4, dtype=np.int32)
    isave = np.zeros(44, dtype=np.int32)
    dsave = np.zeros(29, dtype=dtype)
    trace = []
    xs = []
    calls = 0
    while True:
        ts = task_str(task)
        if not (ts[:2] == "FG" or ts == "NEW_X" or ts == "START"):
            break
        setulb(n, m, x, l, u, nbd, f, g, factr, pgtol, wa, iwa, task, iprint, csave, lsave, isave,
               dsave)
        calls += 1
        ts = task_str(task)
        if ts[:2] == "FG":
            f[0] = fg(x, g)
        elif ts[:5] == "NEW_X":
            row = {
                "iter": int(isave[29]), "nfgv": int(isave[33]), "nseg": int(isave[32]),
                "nact": int(isave[38]), "nfree": int(isave[37]), "nenter": int(isave[40]),
                "nleave": int(n + 1 - isave[39]), "iword": int(isave[36]), "iback": int(isave[24]),
                "col": int(isave[27]), "nskip": int(isave[25]), "nintol": int(isave[21]),
                "stp": float(dsave[13]), "xstep": float(dsave[13] * dsave[3]),
                "theta": float(dsave[0]), "sbgnrm": float(dsave[12]), "f": float(f[0]),
            }
            if want_hash:
                row["hash"], row["hcount"] = setulb.active_set_hash(n, iwa)
            trace.append(row)
            if keep_x:
                xs.append(x.copy())
            if stop is not None:
                s = stop(isave, dsave, float(f[0]))
                if s:
                    task[:] = make_task(s)
        if calls >= max_calls:
            break
    setulb.release(isave)
    out = (trace, task_str(task), x, float(f[0]), isave, dsave)
    return out + (xs,) if keep_x else out


def driver2_stop(limit_nfg=99):
    """test/driver2.f90:174-181 (driver3.f90:200-207 uses 900)."""
    def stop(isave, dsave, f):
        s = None
        if isave[33] >= limit_nfg:
            s = "STOP: TOTAL NO. of f AND g EVALUATIONS EXCEEDS LIMIT"
        if dsave[12] <= 1.0e-10 * (1.0 + abs(f)):
            s = "STOP: THE PROJECTED GRADIENT IS SUFFICIENTLY SMALL"
        return s
    return stop


def iteration_budget_stop(max_iter):
    def stop(isave, dsave, f):
        if isave[29] >= max_iter:
            return "STOP: ITERATION BUDGET"
        return None
    return stop


def fortran_d(v, width, digits):
    """Format like Fortran's `1p,dW.D` edit descriptor (e.g. 1p,d12.5 -> ' 3.46000D+03')."""
    s = "%.*E" % (digits, v)
    mant, exp = s.split("E")
    s = mant + "D" + exp[0] + exp[1:].zfill(2)
    return s.rjust(width)
