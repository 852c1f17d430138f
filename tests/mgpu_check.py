"""Multi-GPU parity check, launched under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/mgpu_check.py [n_global] [m] [l_odd] [iterations] [rosenbrock|quadratic] [x0] [ties]

Runs the sample problem sharded over the N ranks through lbfgsb_setulb_dev_f64 and, on rank 0,
the same problem on one GPU; the per-iterate discrete trace (iter, nfgv, nseg, nfree, nact, iword,
iback, active-set hash) must be equal and f, |proj g| agree to rounding.  Prints one line
`MGPU_CHECK OK ...` or `MGPU_CHECK FAIL ...`; exit code 0 / 1.  With `ties` the sharded run must also have replayed at least
one exit inside a group of equal breakpoints in the reference's heap order (lbfgsb_dev_tie_stats).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import lbfgsb_b200  # noqa: E402
from lbfgsb_b200 import sharded  # noqa: E402


def solve(n_local, off, n_global, m, l_odd, iters, shard, fg_factory, dev, rank, world, kind="rosenbrock", x0=3.0):
    if kind == "quadratic":      # BASELINE.json configs[3]: box [0, 0.5], start in the middle
        x = torch.full((n_local,), 0.25, dtype=torch.float64, device=dev)
        l = torch.zeros(n_local, dtype=torch.float64, device=dev)
        u = torch.full((n_local,), 0.5, dtype=torch.float64, device=dev)
    else:
        x = torch.full((n_local,), x0, dtype=torch.float64, device=dev)
        l = torch.full((n_local,), -100.0, dtype=torch.float64, device=dev)
        l[(off % 2)::2] = l_odd
        u = torch.full((n_local,), 100.0, dtype=torch.float64, device=dev)
    nbd = torch.full((n_local,), 2, dtype=torch.int32, device=dev)
    g = torch.zeros_like(x)
    prob = lbfgsb_b200.DeviceProblem(n_local, m, np.float64, shard=shard)
    fg = fg_factory()
    if shard is not None and hasattr(fg, "engine") and os.environ.get("MGPU_FG", "peer") == "peer":
        fg.engine = prob      # halo and partial f over the workspace's peer-memory exchange (else: torch collectives)
        fg.bounds = (l, u, nbd)   # ... and the objective kernel forms the line-search sums (k_ls_trial is skipped)
    rows = []
    while True:
        prob.setulb_dev(x, l, u, nbd, g, 0.0, 0.0)
        t = prob.task_str()
        if t[:2] == "FG":
            prob.f[0] = fg(x, g)
        elif t[:5] == "NEW_X":
            h, c = prob.active_set_hash()
            if world > 1 and shard is not None:
                ht = torch.tensor([np.int64(np.uint64(h).astype(np.int64)), c], dtype=torch.int64, device=dev)
                dist.all_reduce(ht)
                h, c = int(ht[0]) & 0xFFFFFFFFFFFFFFFF, int(ht[1])
            rows.append(dict(iter=int(prob.isave[29]), nfgv=int(prob.isave[33]), nseg=int(prob.isave[32]),
                             nfree=int(prob.isave[37]), nact=int(prob.isave[38]), iword=int(prob.isave[36]),
                             iback=int(prob.isave[24]), nenter=int(prob.isave[40]), col=int(prob.isave[27]),
                             hash=h, hcount=c, f=float(prob.f[0]), sbgnrm=float(prob.dsave[12]),
                             stp=float(prob.dsave[13]), theta=float(prob.dsave[0])))
            if prob.isave[29] >= iters:
                break
        else:
            break
    task = prob.task_str()
    ties = prob.tie_stats()
    prob.close()
    return rows, task, x, ties


def compare(rows, task, ref, rtask):
    """Sharded trace `rows` against the single-GPU trace `ref`: every discrete field and the active-set hash equal at
    every iterate, f within 1e-10 (first 10 iterates) / 1e-6.  Returns (ok, message, worst relative difference of f)."""
    ok, msg, worst = True, "", 0.0
    if task != rtask or len(rows) != len(ref):
        ok, msg = False, "task/len %r %r %d %d" % (task, rtask, len(rows), len(ref))
    for a, b in zip(rows, ref):
        for k in ("iter", "nfgv", "nseg", "nfree", "nact", "iword", "iback", "nenter", "col", "hash", "hcount"):
            if a[k] != b[k]:
                ok, msg = False, msg + " | it %d %s: %r != %r" % (b["iter"], k, a[k], b[k])
        rel = abs(a["f"] - b["f"]) / max(abs(b["f"]), 1e-300)
        worst = max(worst, rel)
        tol = 1e-10 if b["iter"] <= 10 else 1e-6
        if rel > tol:
            ok, msg = False, msg + " | it %d f rel %.2e" % (b["iter"], rel)
        if not ok:
            break
    return ok, msg, worst


def main():
    n_global = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    l_odd = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    kind = sys.argv[5] if len(sys.argv) > 5 else "rosenbrock"
    x0 = float(sys.argv[6]) if len(sys.argv) > 6 else 3.0
    want_ties = len(sys.argv) > 7 and sys.argv[7] == "ties"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lo, hi = sharded.shard_bounds(n_global, rank, world)
    comm = sharded.nccl_comm_for_engine(rank, world, dist, dev)
    if kind == "quadratic":
        kern = lbfgsb_b200.QuadraticDevice(np.float64, seed=0)
        mk_sharded = lambda: sharded.ShardedQuadraticDevice(kern, lo, rank, world, dist, dev)   # noqa: E731
        mk_single = lambda: (lambda xx, gg: kern(xx, gg, offset=0))                             # noqa: E731
    else:
        kern = lbfgsb_b200.RosenbrockDevice(np.float64)
        mk_sharded = lambda: sharded.ShardedRosenbrockDevice(kern, rank, world, dist, dev)      # noqa: E731
        mk_single = lambda: kern                                                                # noqa: E731
    rows, task, x, ties = solve(hi - lo, lo, n_global, m, l_odd, iters, (lo, n_global, comm, rank, world),
                                mk_sharded, dev, rank, world, kind, x0)
    ok = True
    if rank == 0:
        ref, rtask, xr, rties = solve(n_global, 0, n_global, m, l_odd, iters, None, mk_single, dev, 0, 1, kind, x0)
        ok, msg, worst = compare(rows, task, ref, rtask)
        if want_ties and not (ties[0] >= 1 and ties[1] == 0 and rties[0] >= 1):
            ok, msg = False, msg + " | tie replays sharded %r single %r" % (ties, rties)
        walks = [r["nseg"] for r in ref if r["nseg"] > 1]
        print("MGPU_CHECK %s %s world=%d n=%d m=%d l_odd=%g x0=%g iterations=%d walks(nseg>1)=%s tie_replays(sharded,single)=%d,%d worst_rel_f=%.2e %s" % (
            "OK" if ok else "FAIL", kind, world, n_global, m, l_odd, x0, len(ref), walks[:6], ties[0], rties[0], worst, msg[:600]), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    lbfgsb_b200.lib().lbfgsb_dev_nccl_destroy(comm)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
