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


_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(v):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic), as lb_mix64 in csrc/engine.cu."""
    with np.errstate(over="ignore"):
        v = v + np.uint64(0x9E3779B97F4A7C15)
        v = (v ^ (v >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        v = (v ^ (v >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return v ^ (v >> np.uint64(31))


def quadratic_coefficients(lo, hi, seed=0, dtype=np.float64):
    """diag (2 + delta_i) and b_i of the convex quadratic of BASELINE.json configs[3] for the global
    indices [lo, hi) (numpy reference of quad_coeff in csrc/engine.cu)."""
    with np.errstate(over="ignore"):
        i = np.arange(lo, hi, dtype=np.uint64)
        sp = np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        u1 = (_mix64(np.uint64(2) * i + np.uint64(2) * sp) >> np.uint64(32)).astype(np.float64) * (1.0 / 4294967296.0)
        u2 = (_mix64(np.uint64(2) * i + np.uint64(1) + np.uint64(2) * sp) >> np.uint64(32)).astype(np.float64) * (1.0 / 4294967296.0)
    return (2.0 + (0.1 + u1)).astype(dtype), (2.0 * u2 - 1.0).astype(dtype)


def quadratic_shard_numpy(x, lo, seed, xl, xr):
    """f contribution and gradient of one shard of f = 1/2 x'Ax - b'x, A = tridiag(-1, 2+delta, -1)
    (numpy reference of lbfgsb_problem_quadratic_*; same operation order per element).
    Returns (this shard's part of f, g)."""
    n = x.shape[0]
    diag, b = quadratic_coefficients(lo, lo + n, seed, x.dtype)
    xprev = np.empty_like(x)
    xprev[1:] = x[:-1]
    xprev[0] = xl
    xnext = np.empty_like(x)
    xnext[:-1] = x[1:]
    xnext[-1] = xr
    ax = diag * x - xprev - xnext
    g = ax - b
    half = x.dtype.type(0.5)
    return float(((half * ax - b) * x).sum(dtype=np.float64)), g


def quadratic_problem(n_local, dtype=np.float64, scale=0.5):
    """Box [0, scale] on every variable (nbd = 2), start in the middle of the box.  With scale = 0.5 about
    half of the variables end on a bound (measured with the oracle at n = 2e4: 36.5% on l = 0, 13.7% on
    u = 0.5; the unconstrained minimiser has |x| up to 2.3 and half of its components negative)."""
    l = np.zeros(n_local, dtype=dtype)
    u = np.full(n_local, scale, dtype=dtype)
    nbd = np.full(n_local, 2, dtype=np.int32)
    x = np.full(n_local, 0.5 * scale, dtype=dtype)
    return x, l, u, nbd


def _halo_on_device(x, rank, world, dist):
    """(halo, fpart) device tensors: halo = [last value of the left neighbour, first value of the right one]
    (0 at the ends of the chain) from one all-gather of the two edge values, never leaving the device."""
    import torch
    edge = torch.stack([x[0], x[-1]])
    allv = torch.empty(2 * world, dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(allv, edge)
    zero = torch.zeros((), dtype=x.dtype, device=x.device)
    xl = allv[2 * (rank - 1) + 1] if rank > 0 else zero
    xr = allv[2 * (rank + 1)] if rank < world - 1 else zero
    return torch.stack([xl, xr]), torch.empty(1, dtype=x.dtype, device=x.device)


class ShardedQuadraticDevice:
    """Device f/g of the convex quadratic on a shard: halo exchange + partial-f all-reduce around
    lbfgsb_problem_quadratic_* (lbfgsb_b200.QuadraticDevice)."""

    def __init__(self, kernel, lo, rank, world, dist, device, engine=None):
        self.k, self.lo, self.rank, self.world, self.dist, self.device = kernel, lo, rank, world, dist, device
        self.engine = engine     # a sharded DeviceProblem: its peer-memory exchange carries halo and partial f
        self.bounds = None       # (l, u, nbd) device tensors: the kernel then also forms the line-search sums

    def __call__(self, x, g):
        import torch
        if self.world == 1:
            return self.k(x, g, offset=self.lo)
        if self.engine is not None and x.is_cuda and x.dtype == torch.float64:
            f = self.engine.sharded_fg(1, x, g, *(self.bounds or (None, None, None)), seed=self.k.seed)
            if f is not None:
                return f
        if x.is_cuda and x.dtype == torch.float64:
            halo, fpart = _halo_on_device(x, self.rank, self.world, self.dist)
            self.k.shard_async(x, g, self.lo, halo, fpart)
            self.dist.all_reduce(fpart)
            return float(fpart)      # the only host round trip of the evaluation
        edge = torch.stack([x[0], x[-1]])
        allv = [torch.empty_like(edge) for _ in range(self.world)]
        self.dist.all_gather(allv, edge)
        r = self.rank
        xl = float(allv[r - 1][1]) if r > 0 else 0.0
        xr = float(allv[r + 1][0]) if r < self.world - 1 else 0.0
        fl = self.k(x, g, offset=self.lo, xl=xl, xr=xr)
        ft = torch.tensor([fl], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(ft)
        return float(ft)


class ShardedRosenbrockDevice:
    """Device f/g of the sample problem on a shard: halo exchange + partial-f all-reduce around
    lbfgsb_problem_rosenbrock_* (lbfgsb_b200.RosenbrockDevice)."""

    def __init__(self, kernel, rank, world, dist, device, engine=None):
        self.k, self.rank, self.world, self.dist, self.device = kernel, rank, world, dist, device
        self.engine = engine     # a sharded DeviceProblem: its peer-memory exchange carries halo and partial f
        self.bounds = None       # (l, u, nbd) device tensors: the kernel then also forms the line-search sums

    def __call__(self, x, g):
        import torch
        if self.world == 1:
            return self.k(x, g)
        if self.engine is not None and x.is_cuda and x.dtype == torch.float64:
            f = self.engine.sharded_fg(0, x, g, *(self.bounds or (None, None, None)))
            if f is not None:
                return f
        if x.is_cuda and x.dtype == torch.float64:
            halo, fpart = _halo_on_device(x, self.rank, self.world, self.dist)
            self.k.shard_async(x, g, 1 if self.rank == 0 else 0, 1 if self.rank == self.world - 1 else 0, halo, fpart)
            self.dist.all_reduce(fpart)
            return float(fpart)      # the only host round trip of the evaluation
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


class HostShard:
    """Host-buffer front of one rank's shard of a sharded problem: the staging that the C-ABI host twin
    (lbfgsb_setulb_f64, single GPU) does around the device variant, done here around lbfgsb_setulb_dev_* of a sharded
    workspace.  The caller keeps x, l, u, nbd, g of its shard in (preferably pinned) host arrays and drives the same
    `task` protocol (src/lbfgsb.f90:88-89); per call g goes host -> device on an 'FG' re-entry and x comes back
    whenever the call moved it."""

    def __init__(self, n_local, lo, n_global, m, comm, rank, world, dtype=np.float64):
        import torch
        import lbfgsb_b200
        self.torch = torch
        tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        self.prob = lbfgsb_b200.DeviceProblem(n_local, m, dtype, shard=(lo, n_global, comm, rank, world))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.xd, self.ld, self.ud, self.gd = (torch.empty(n_local, dtype=tdt, device=dev) for _ in range(4))
        self.nd = torch.empty(n_local, dtype=torch.int32, device=dev)

    def setulb(self, x, l, u, nbd, f, g, factr, pgtol, task, iprint, csave, lsave, isave, dsave):
        torch = self.torch
        p = self.prob
        t_in = bytes(task[:5])
        if t_in == b"START":
            for d, h in ((self.xd, x), (self.ld, l), (self.ud, u), (self.nd, nbd)):
                d.copy_(torch.from_numpy(h), non_blocking=True)
        elif t_in[:2] == b"FG":
            self.gd.copy_(torch.from_numpy(g), non_blocking=True)
        torch.cuda.synchronize()
        p.task, p.csave, p.lsave, p.isave, p.dsave, p.f = task, csave, lsave, isave, dsave, f
        p.setulb_dev(self.xd, self.ld, self.ud, self.nd, self.gd, factr, pgtol)
        t_out = bytes(task[:5])
        if t_in == b"START" or t_out[:2] == b"FG":
            torch.from_numpy(x).copy_(self.xd, non_blocking=True)          # projected start / next trial point
        elif t_out != b"NEW_X":
            torch.from_numpy(x).copy_(self.xd, non_blocking=True)          # termination: x (and g) may have been restored
            torch.from_numpy(g).copy_(self.gd, non_blocking=True)
        torch.cuda.synchronize()

    def close(self):
        self.prob.close()


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
