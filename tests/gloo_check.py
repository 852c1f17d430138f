"""World-size-2 (or more) CPU check of the host-side sharding logic, run under torchrun with the
gloo backend by tests/test_sharded_cpu.py:
  - shard_bounds partitions [0, n) into contiguous blocks,
  - halo exchange + partial-f all-reduce reproduce the global f and g of the sample problem
    (test/driver1.f90:274-289) and of the convex quadratic of BASELINE.json configs[3] exactly as
    the unsharded routines compute them.
"""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from lbfgsb_b200 import sharded  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for n in (7, 1000, 12345):
        lo, hi = sharded.shard_bounds(n, rank, world)
        spans = [None] * world
        dist.all_gather_object(spans, (lo, hi))
        ok &= spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        ok &= max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
        rng = np.random.default_rng(n)
        xg = rng.uniform(-2, 2, n)
        x = xg[lo:hi].copy()
        xl, xr = sharded.exchange_halo(x[0], x[-1], rank, world, dist)
        fpart, g = sharded.rosenbrock_shard_numpy(x, rank == 0, rank == world - 1, xl, xr)
        f = sharded.allreduce_sum(fpart, dist)
        fref, gref = sharded.rosenbrock_shard_numpy(xg, True, True, 0.0, 0.0)
        ok &= abs(f - fref) <= 1e-12 * abs(fref)
        ok &= np.array_equal(g, gref[lo:hi])
        # the convex quadratic of BASELINE.json configs[3]: hashed coefficients by global index
        qpart, qg = sharded.quadratic_shard_numpy(x, lo, 11, xl, xr)
        qf = sharded.allreduce_sum(qpart, dist)
        qfref, qgref = sharded.quadratic_shard_numpy(xg, 0, 11, 0.0, 0.0)
        ok &= abs(qf - qfref) <= 1e-12 * max(1.0, abs(qfref))
        ok &= np.array_equal(qg, qgref[lo:hi])
        if rank > 0:
            ok &= xl == xg[lo - 1]
        if rank < world - 1:
            ok &= xr == xg[hi]
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        print("GLOO_CHECK", "OK" if all(flags) else "FAIL", "world=%d" % world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
