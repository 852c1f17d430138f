"""Every launch of ONE iteration, in order, from an ncu launch list (gpu__time_duration.sum pass).

    python tools/iteration_launches.py LAUNCHES_CSV OUT_MD "title" [which]

The launches between the `which`-th and the next `k_update_classify` launch (default: the second) are one iteration
(NEW_X entry to NEW_X entry, the objective and the line-search entry included).  Per-launch times under ncu are cold-cache
and serialised: compare shares, not absolute values."""
import csv
import re
import sys
from collections import OrderedDict

src, out, title = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr, seq = None, []
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
    sc = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(d.get("Metric Unit", "ns"), 1e-3)
    seq.append((re.sub(r"\(.*", "", d["Kernel Name"].replace("void ", "")), v * sc))
idx = [i for i, (n, _) in enumerate(seq) if n.startswith("k_update_classify")]
a, b = idx[which], idx[which + 1]
agg, tot = OrderedDict(), 0.0
for n, t in seq[a:b]:
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += t
    tot += t
with open(out, "w") as fh:
    fh.write("# %s\n\n" % title)
    fh.write("The launches between two consecutive `k_update_classify` launches of the ncu launch list (`--metrics "
             "gpu__time_duration.sum --clock-control none`);\nper-launch times are cold-cache and serialised: compare shares.\n\n")
    fh.write("| kernel (in order of first launch) | launches | total us | share |\n|---|---:|---:|---:|\n")
    for n, (c, t) in agg.items():
        fh.write("| `%s` | %d | %.1f | %.1f%% |\n" % (n, c, t, 100 * t / tot))
    fh.write("\nTotal %.0f us over %d launches.\n" % (tot, b - a))
print("wrote", out, b - a, "launches", round(tot), "us")
