"""Development aid: time per iteration where the fixed cost per setulb call decides it (small n, or the per-GPU share of a
strongly scaled problem) -- ms per iteration, launches and host read-backs per iteration, and the scalar kernels' time.
  python tools/small_n_latency.py [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import lbfgsb_b200  # noqa: E402


def run(n, m, iters=60, warm=20, profile=False):
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream()
    x = torch.full((n,), 3.0, dtype=torch.float64, device=dev)
    l = torch.full((n,), -100.0, dtype=torch.float64, device=dev); l[0::2] = 1.1
    u = torch.full((n,), 100.0, dtype=torch.float64, device=dev)
    nbd = torch.full((n,), 2, dtype=torch.int32, device=dev)
    g = torch.zeros_like(x)
    torch.cuda.synchronize()
    prob = lbfgsb_b200.DeviceProblem(n, m, np.float64, stream=st.cuda_stream)
    t0 = l0 = s0 = None
    nfg = 0
    while True:
        prob.setulb_dev(x, l, u, nbd, g, 0.0, 0.0)
        t = bytes(prob.task[:5])
        if t[:2] == b"FG":
            prob.f[0] = prob.fused_fg(0, x, g, l, u, nbd); nfg += 1
        elif t == b"NEW_X":
            it = int(prob.isave[29])
            if it == warm:
                torch.cuda.synchronize()
                if profile:
                    prob.profile(True); prob.profile_reset()
                t0 = time.perf_counter(); (l0, s0) = prob.counters(); f0 = nfg
            if it >= warm + iters:
                break
        else:
            break
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    l1, s1 = prob.counters()
    k = int(prob.isave[29]) - warm
    out = {"n": n, "m": m, "iterations": k, "ms_per_iteration": dt / k * 1e3, "launches_per_iteration": (l1 - l0 + 2 * (nfg - f0)) / k,
           "host_readbacks_per_iteration": (s1 - s0 + (nfg - f0)) / k, "fg_per_iteration": (nfg - f0) / k, "task": prob.task_str()}
    if profile:
        pr = prob.profile_read()
        out["families_us_per_call"] = {a: (round(v["ms"] / v["calls"] * 1e3, 1), v["calls"]) for a, v in pr.items() if v["calls"]}
    prob.close()
    return out


def run_graph(n, m, iters=60, warm=20):
    """The same iterations through lbfgsb_minimize_graph_dev_f64: one CUDA-graph launch and one read-back per step."""
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream()
    x = torch.full((n,), 3.0, dtype=torch.float64, device=dev)
    l = torch.full((n,), -100.0, dtype=torch.float64, device=dev); l[0::2] = 1.1
    u = torch.full((n,), 100.0, dtype=torch.float64, device=dev)
    nbd = torch.full((n,), 2, dtype=torch.int32, device=dev)
    g = torch.zeros_like(x)
    halo = torch.zeros(2, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    fg = lbfgsb_b200.RosenbrockDevice(np.float64, stream=st.cuda_stream); fg._n = n
    times = []
    for budget in (warm, warm + iters):      # two solves of the same problem: the difference is `iters` iterations
        x.fill_(3.0)
        prob = lbfgsb_b200.DeviceProblem(n, m, np.float64, stream=st.cuda_stream)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        prob.minimize_graph(x, l, u, nbd, g, fg.enqueue(halo), 0.0, 0.0, max_iter=budget)
        torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
        steps, per = prob.graph_stats(); l1, s1 = prob.counters(); nfgv = int(prob.isave[33]); it = int(prob.isave[29])
        prob.close()
    return {"n": n, "m": m, "mode": "minimize_graph", "iterations": iters, "ms_per_iteration": (times[1] - times[0]) / iters * 1e3,
            "graph_steps_of_the_longer_solve": steps, "engine_kernels_per_graph_launch": per, "iterations_of_the_longer_solve": it,
            "fg_evaluations_of_the_longer_solve": nfgv}


def main():
    res = []
    for n in (1000, 100000, 1000000, 12500000):
        res.append(run(n, 10))
        print(json.dumps(res[-1]), flush=True)
        res.append(run_graph(n, 10))
        print(json.dumps(res[-1]), flush=True)
    res.append(run(12500000, 10, profile=True))
    print(json.dumps(res[-1]), flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
