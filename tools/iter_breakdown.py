"""Development aid: where one iteration's time goes, iteration by iteration.

    python tools/iter_breakdown.py [rosenbrock|quadratic] [n] [iterations]

Per iteration after the warm-up: wall time (host clock around the whole iteration, GPU idle included), the time of the
objective calls (host clock, includes the read-back of f), the engine's own phase timers dsave(7:9) (CUDA events on its
stream: cauchy, subspace, line search), nseg / entering / leaving counts, launches and host read-backs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench

kind = sys.argv[1] if len(sys.argv) > 1 else "quadratic"
wl = bench.WORKLOADS[kind]
n = int(sys.argv[2]) if len(sys.argv) > 2 else wl["n"]
K = int(sys.argv[3]) if len(sys.argv) > 3 else 12
cx = bench.Ctx()
torch = cx.torch
ds = bench.DeviceSolve(cx, kind, n, wl["m"], wl["dtype"], bench.L_ODD)
p = ds.prob
with torch.cuda.stream(cx.stream):
    ds.step_to(wl["m"] + 4)
    torch.cuda.synchronize()
    fg_t = [0.0]
    inner = ds._fg

    def timed_fg():
        t0 = time.perf_counter(); v = inner(); fg_t[0] += time.perf_counter() - t0
        return v
    ds._fg = timed_fg
    print("%s n=%d m=%d: iter  wall ms | fg ms (calls) | cauchy  subspace  lnsrch (engine events, ms) | setulb host ms | nseg nenter nleave | launches syncs" % (kind, n, wl["m"]))
    for k in range(K):
        it0 = ds.counts()["iter"]
        d0 = [float(p.dsave[6]), float(p.dsave[7]), float(p.dsave[8])]
        l0, s0 = p.counters(); f0 = ds.nfg; fg_t[0] = 0.0
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ds.step_to(it0 + 1)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        c = ds.counts(); l1, s1 = p.counters()
        d1 = [float(p.dsave[6]), float(p.dsave[7]), float(p.dsave[8])]
        wall = (t1 - t0) * 1e3
        ph = [(b - a) * 1e3 for a, b in zip(d0, d1)]
        print("  %3d  %7.2f | %5.2f (%d) | %6.2f  %6.2f  %6.2f | %6.2f | %7d %8d %8d | %3d %2d" % (
            c["iter"], wall, fg_t[0] * 1e3, ds.nfg - f0, ph[0], ph[1], ph[2], wall - fg_t[0] * 1e3, c["nseg"], c["nenter"], c["nleave"], l1 - l0, s1 - s0))
