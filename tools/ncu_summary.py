"""Turns the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/ncu_summary.py ROUND_TAG LAUNCHES_CSV FULL_REP N [KEY_SUFFIX]

FULL_REP is a .ncu-rep, or the csv that `ncu -i rep --page raw --csv` wrote on the GPU box (the reports of n = 4e8 runs are
too large to bring back).  KEY_SUFFIX (e.g. "@f32m20") keys the traffic table for captures of other instantiations than
real64 / m <= 10; launches that returned at once (< 0.05 ms) are left out.

  profiles/<tag>_launches.md     per-kernel share of the profiled command (gpu__time_duration pass)
  profiles/<tag>_ncu_full.md     the --set full metrics that matter for an HBM-bound kernel
  profiles/ncu_traffic.json      kernel family -> DRAM bytes per launch (bench.py's roofline.traffic)
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAM = {"k_update": "update", "k_cauchy_classify": "cauchy_classify", "k_formk_gram": "formk_gram",
       "k_cmprlb_wv": "cmprlb_wv", "k_subsm_step": "subsm_step", "k_ls_init": "ls_init", "k_ls_trial": "ls_trial",
       "k_ls_step": "ls_step", "k_gcp_freev": "gcp_freev", "k_iter_head": "iter_head",
       "k_update_classify": "update_classify", "k_formk_cmprlb": "formk_cmprlb", "k_subsm_lsinit": "subsm_lsinit",
       "k_formk_delta": "formk_delta"}


def short(name):
    n = name.split("(")[0].replace("void ", "").strip()
    return n


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    per = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
        k = short(d["Kernel Name"])
        per[k][0] += 1
        per[k][1] += v * scale
    tot = sum(v[1] for v in per.values())
    with open(out, "w") as fh:
        fh.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, ms) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, c, ms, 100 * ms / tot))
        fh.write("\nTotal GPU time of the profiled command: %.1f ms over %d launches (cold-cache, serialised: compare shares).\n" % (tot, sum(v[0] for v in per.values())))


def full(rep, out, n, suffix=""):
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:  # noqa: BLE001
        pass
    with open(out, "w") as fh:
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = short(d["Kernel Name"])
            try:
                if float(d["gpu__time_duration.sum"].replace(",", "")) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u["gpu__time_duration.sum"], 1.0) < 0.05:
                    continue
            except Exception:  # noqa: BLE001
                pass
            fh.write("## `%s`\n\n| metric | value |\n|---|---|\n" % d["Kernel Name"].strip())
            for w in want:
                if w in d:
                    fh.write("| %s | %s %s |\n" % (w, d[w], u[w]))
            stalls = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(d[h].replace(",", ""))) for h in hdr
                      if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and d[h] not in ("", "n/a")]
            if not stalls:
                stalls = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(d[h].replace(",", "")))
                          for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")]
            tot = sum(v for _, v in stalls) or 1.0
            stalls.sort(key=lambda kv: -kv[1])
            fh.write("| top stall reasons | %s |\n\n" % ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in stalls[:5]))

            def gb(x, un):
                v = float(x.replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(un, 1.0)
            base = name.split("<")[0]
            if base in FAM and "dram__bytes_read.sum" in d:
                traffic[FAM[base] + suffix] = {"n": int(n), "kernel": d["Kernel Name"].strip(),
                                      "dram_bytes_per_launch": gb(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) +
                                      gb(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]),
                                      "duration_ms_under_ncu": d["gpu__time_duration.sum"] + " " + u["gpu__time_duration.sum"]}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as fh:
        json.dump(traffic, fh, indent=1)


if __name__ == "__main__":
    tag, lcsv, rep, n = sys.argv[1:5]
    suffix = sys.argv[5] if len(sys.argv) > 5 else ""
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if os.path.exists(lcsv):
        launches(lcsv, os.path.join(ROOT, "profiles", tag + "_launches.md"))
    if os.path.exists(rep):
        full(rep, os.path.join(ROOT, "profiles", tag + "_ncu_full.md"), n, suffix)
