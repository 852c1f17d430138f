"""Opcode counts per kernel of the built library -> profiles/sass_summary.md.

    python tools/sass_summary.py [lbfgsb_b200/liblbfgsb_b200.so]

Runs `cuobjdump -sass` on the sm_100a cubin inside the shared library (no GPU needed) and counts, per kernel, the
mnemonics that show how the kernel moves data: UBLKCP (1-D bulk copy, the TMA engine), SYNCS (mbarrier operations),
LDS.128 / LDG.E.128 / STG.E.128 (128-bit shared / global accesses), DADD/DMUL/DFMA and FADD/FMUL/FFMA (the engine is
built with -fmad=false, so the fused forms should be absent from the arithmetic of the passes), BAR.SYNC, plus the
total instruction count."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lbfgsb_b200", "liblbfgsb_b200.so")
out = os.path.join(ROOT, "profiles", "sass_summary.md")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = {}
names = re.findall(r"Function : (\S+)", txt)
if names:
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, dm))
COLS = ["UBLKCP", "SYNCS", "LDS.128", "LDS.64", "LDG.E.128", "LDG.E.64", "STG.E.128", "LDG", "STG", "DADD", "DMUL", "DFMA", "FADD", "FMUL", "FFMA",
        "BAR.SYNC", "SHFL", "ATOM", "RED"]
rows = []
arch = re.search(r"arch = (sm_\w+)", txt)
cur, cnt, total = None, None, 0


def flush():
    if cur is not None:
        rows.append((cur, total, dict(cnt)))


for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        cur, cnt, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        total += 1
        for c in COLS:
            if op == c or op.startswith(c + "."):
                cnt[c] += 1
flush()


def short(mangled):
    s = demangle.get(mangled, mangled)
    s = re.sub(r"^void ", "", s).strip()
    if s.endswith(")"):   # drop the parameter list: the parenthesis that matches the last one
        depth = 0
        for k in range(len(s) - 1, -1, -1):
            depth += (s[k] == ")") - (s[k] == "(")
            if depth == 0:
                s = s[:k]
                break
    return s.replace("(int)", "").replace("(bool)", "")


rows.sort(key=lambda r: -r[1])
with open(out, "w") as fh:
    fh.write("# SASS opcode summary of `lbfgsb_b200/liblbfgsb_b200.so` (%s)\n\n" % (arch.group(1) if arch else "?"))
    fh.write("Produced by `python tools/sass_summary.py` (`cuobjdump -sass`, no GPU). `LDG`/`STG` count every width, the `.128`\n"
             "columns the 128-bit forms among them. `UBLKCP` = `cp.async.bulk` (1-D TMA bulk copy), `SYNCS` = mbarrier\n"
             "arrive/try_wait/expect_tx. The engine is compiled with `-fmad=false`: `DFMA`/`FFMA` that remain are address\n"
             "arithmetic or the explicit fma() of the objective kernels, not contractions of the passes' sums.\n\n")
    fh.write("| kernel | instr | " + " | ".join(COLS) + " |\n|---|---:|" + "---:|" * len(COLS) + "\n")
    for name, tot, c in rows:
        fh.write("| `%s` | %d | %s |\n" % (short(name), tot, " | ".join(str(c.get(k, 0)) for k in COLS)))
    fh.write("\n%d kernels; %d use UBLKCP (the TMA-staged S/Y passes), tcgen05/UTCMMA: %d (none by design: every contraction is n x 2m, HBM-bound).\n"
             % (len(rows), sum(1 for r in rows if r[2].get("UBLKCP")), len(re.findall(r"UTC\w*MMA", txt))))
print("wrote", out, len(rows), "kernels")
