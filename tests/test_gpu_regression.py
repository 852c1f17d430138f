"""Bit-level regression of the engine's per-iterate traces (tools/trace_digest.py).

The reductions have a fixed shape (include/lbfgsb_b200_shape.h), so every iterate of a given problem
is bit-reproducible: across runs, across GPUs, and across re-organisations of the kernels that keep the
shape -- the fused passes (k_update_classify, k_formk_cmprlb) must give exactly the bits of the
separate passes they replace.  tests/golden/gpu_trace_digest.json was recorded on a B200 (first with the
unfused kernels; re-recorded for the two cases whose Cauchy search crosses a round boundary when the
breakpoint walk was split into rounds, which re-associates the prefix sums there -- every case still
passes the oracle parity tests; and once more for n100001_m5 when formk's entering/leaving corrections got a
register-tiled kernel, which adds the listed rows in a different order; the two REAL32 cases in round 2, when the REAL32
shape became VEC = 2 / UNROLL = 8 -- same tiles, another thread -> element map inside a tile -- with all REAL64 digests
unchanged by that build; and the three cases with entering/leaving rows again when k_formk_delta got
sorted, double-buffered tiles and a one-wave grid -- other grouping of the listed rows); it is compared here with a fresh run, with the fused passes on and off.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "gpu_trace_digest.json")


def _run(env_extra, tmp_path, tag):
    out = os.path.join(str(tmp_path), "digest_%s.json" % tag)
    env = dict(os.environ)
    env.update(env_extra)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "trace_digest.py"), out], env=env,
                          stdout=subprocess.DEVNULL)
    return json.load(open(out))


@pytest.mark.gpu
@pytest.mark.parametrize("fusion", ["fused", "unfused", "fused+fg_epilogue", "fused_general_pipeline", "fused_no_bp_hint"])
def test_trace_bits_match_recorded(fusion, tmp_path):
    """fused: the fast NEW_X pipeline (merged scalar kernels, pause/resume); fused_general_pipeline: the same fused passes
    driven by the general pipeline (LBFGSB_B200_NO_FAST=1); fused_no_bp_hint: cauchy's per-variable pass never stores the
    breakpoints ahead of a walk (the walk's first pass computes them).  fused+fg_epilogue: the objective kernel forms the line-search sums itself (lbfgsb_problem_fused_f64) and the
    engine's k_ls_trial is skipped -- same products in the same order, hence the same bits."""
    gold = json.load(open(GOLD))
    got = _run({"LBFGSB_B200_NO_FUSION": "1" if fusion == "unfused" else "0",
                "LBFGSB_B200_NO_FAST": "1" if fusion == "fused_general_pipeline" else "0",
                "LBFGSB_B200_NO_BP_HINT": "1" if fusion == "fused_no_bp_hint" else "0",
                "LBFGSB_DIGEST_FUSED_FG": "1" if fusion.endswith("fg_epilogue") else "0"}, tmp_path, fusion.replace("+", "_"))
    assert set(got) == set(gold)
    for name in sorted(gold):
        for key in ("iterations", "task", "sha256"):
            assert got[name][key] == gold[name][key], (fusion, name, key, got[name], gold[name])
