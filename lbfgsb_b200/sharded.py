"""Host-side plumbing of a problem sharded over the GPUs of one node (one process per GPU).

The engine shards by contiguous blocks of variables; its own reductions
travel through NCCL inside the C++ engine.  What the *caller* has to do per f/g evaluation of a
chain-coupled objective such as the reference's sample problem (test/driver1.f90:274-289) is
  - exchange one boundary value of x with each neighbour (halo),
  - sum the partial f over the ranks,
and that is what this module holds, on top of torch.distributed (nccl on GPUs, gloo in the CPU
tests).  No collective is used for anything else.
"""
import numpy as np


def shard_bounds(n_global, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; the first n_global % world ranks get one more."""
    base, rem = divmod(int(n_global), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exchange_halo(first, last, rank, world, dist, device="cpu"):
    """Returns (xl, xr): the last value of the left neighbour and the first value of the right one
    (0.0 at the ends).  One all_gather of 2 numbers per rank."""
    import torch
    edge = torch.tensor([float(first), float(last)], dtype=torch.float64, device=device)
    allv = [torch.empty_like(edge) for _ in range(world)]
    dist.all_gather(allv, edge)
    xl = float(allv[rank - 1][1]) if rank > 0 else 0.0
    xr = float(allv[rank + 1][0]) if rank < world - 1 else 0.0
    return xl, xr


def allreduce_sum(value, dist, device="cpu"):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t)
    return float(t)


def rosenbrock_shard_numpy(x, first, last, xl, xr):
    """f contribution and gradient of one shard of the extended Rosenbrock function of
    test/driver1.f90:274-289 (numpy reference of lbfgsb_problem_rosenbrock_*):
        f = 4 [ 1/4 (x_1 - 1)^2 + sum_{i>=2} (x_i - x_{i-1}^2)^2 ].
    Returns (this shard's part of f, g); the parts of all shards add up to f."""
    n = x.shape[0]
    xprev = np.empty(n)
    xprev[1:] = x[:-1]
    xprev[0] = xl
    xnext = np.empty(n)
    xnext[:-1] = x[1:]
    xnext[-1] = xr
    t2 = x - xprev ** 2          # x(i) - x(i-1)^2
    t1 = xnext - x ** 2          # x(i+1) - x(i)^2
    g = 8.0 * t2 - 16.0 * x * t1
    terms = t2 ** 2
    if last:
        g[-1] = 8.0 * t2[-1]
    if first:
        g[0] = 2.0 * (x[0] - 1.0) - 16.0 * x[0] * t1[0]
        terms[0] = 0.25 * (x[0] - 1.0) ** 2
    return 4.0 * float(terms.sum()), g


class ShardedRosenbrockDevice:
    """Device f/g of the sample problem on a shard: halo exchange + partial-f all-reduce around
    lbfgsb_problem_rosenbrock_* (lbfgsb_b200.RosenbrockDevice)."""

    def __init__(self, kernel, rank, world, dist, device):
        self.k, self.rank, self.world, self.dist, self.device = kernel, rank, world, dist, device

    def __call__(self, x, g):
        import torch
        if self.world == 1:
            return self.k(x, g)
        edge = torch.stack([x[0], x[-1]])
        allv = [torch.empty_like(edge) for _ in range(self.world)]
        self.dist.all_gather(allv, edge)
        r = self.rank
        xl = float(allv[r - 1][1]) if r > 0 else 0.0
        xr = float(allv[r + 1][0]) if r < self.world - 1 else 0.0
        fl = self.k(x, g, first=1 if r == 0 else 0, last=1 if r == self.world - 1 else 0, xl=xl, xr=xr)
        ft = torch.tensor([fl], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(ft)
        return float(ft)


def nccl_comm_for_engine(rank, world, dist, device):
    """Creates the engine's own NCCL communicator: the 128-byte unique id is made on rank 0 and
    broadcast through torch.distributed."""
    import ctypes as C
    import torch
    import lbfgsb_b200
    L = lbfgsb_b200.lib()
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (C.c_char * 128)()
        if L.lbfgsb_dev_nccl_unique_id(raw) != 0:
            raise lbfgsb_b200.LbfgsbB200Error("nccl unique id: " + lbfgsb_b200.last_error())
        idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    idd = idbuf.to(device)
    dist.broadcast(idd, 0)
    comm = L.lbfgsb_dev_nccl_init(idd.cpu().numpy().tobytes(), rank, world)
    if not comm:
        raise lbfgsb_b200.LbfgsbB200Error("nccl init: " + lbfgsb_b200.last_error())
    return comm
