"""Sharded (N > 1 GPUs) vs single-GPU parity; skipped on a box with one GPU.
Run on a multi-GPU box:  gpurun --gpus 2 -- python -m pytest tests/test_gpu_mgpu.py -m gpu"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("args,fg", [(("200000", "5", "1.0", "30"), "peer"), (("100001", "10", "1.1", "25"), "peer"),
                                     (("4099", "7", "1.5", "20"), "peer"), (("4099", "7", "1.5", "20"), "torch"),
                                     (("150001", "10", "0", "25", "quadratic"), "peer"),
                                     (("150001", "10", "0", "25", "quadratic"), "torch")])
def test_sharded_matches_single_gpu(args, fg):
    """fg = "peer": the sample objective's halo and partial f travel over the workspace's peer-memory exchange
    (lbfgsb_problem_sharded_f64); "torch": through torch.distributed collectives (lbfgsb_problem_*_halo_f64)."""
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ng < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mgpu_check.py")] + list(args)
    env = dict(os.environ)
    env["MGPU_FG"] = fg
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and "MGPU_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("args", [("1000", "5", "2.0", "12", "rosenbrock", "3.0", "ties"), ("3001", "3", "2.5", "12", "rosenbrock", "5.0", "ties"),
                                  ("200000", "10", "1.3", "12", "rosenbrock", "5.0", "ties"), ("4000001", "3", "2.5", "6", "rosenbrock", "5.0", "ties")])
def test_sharded_tie_exit_follows_the_heap_order(args):
    """The problems of tests/test_gpu_rare_paths.py whose Cauchy search ends inside a group of equal breakpoints, sharded:
    every rank gathers the breakpoints of the call, pops the reference's heap (hpsolb, src/lbfgsb.f90:2079-2157) and the
    group runs as a round of its own in that order -- the active set (hash) of every iterate equals the single-GPU run's,
    which the rare-path tests compare with the oracle."""
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ng < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mgpu_check.py")] + list(args)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ))
    assert r.returncode == 0 and "MGPU_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
