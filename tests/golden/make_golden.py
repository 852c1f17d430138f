"""Extracts the reference's known-answer vectors into tests/golden/reference_outputs.json.

Source (read-only, only available in the build container):
  /root/reference/test/OUTPUTS/output_90_{1,2,3}, output_77_{1,2,3}, iterate.dat
Run:  python tests/golden/make_golden.py
The JSON is committed; tests never read /root/reference.
"""
import json
import os
import re

REF = "/root/reference/test/OUTPUTS"
HERE = os.path.dirname(os.path.abspath(__file__))


def d2f(s):
    return float(s.replace("D", "E"))


def parse_driver1(path):
    out = {"iterates": []}
    for line in open(path):
        m = re.match(r"At iterate\s+(\d+)\s+f=\s+(\S+)\s+\|proj g\|=\s+(\S+)", line)
        if m:
            out["iterates"].append({"iter": int(m.group(1)), "f_str": m.group(2), "pg_str": m.group(3)})
        m = re.match(r"\s+F =\s+(\S+)", line)
        if m:
            out["final_f"] = float(m.group(1).replace("E-0", "E-"))
            out["final_f_str"] = m.group(1)
        m = re.match(r"\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\S+D\S+)\s+(\S+D\S+)\s*$", line)
        if m:
            out["summary"] = {"n": int(m.group(1)), "tit": int(m.group(2)), "tnf": int(m.group(3)),
                              "tnint": int(m.group(4)), "skip": int(m.group(5)), "nact": int(m.group(6)),
                              "projg_str": m.group(7), "f_str": m.group(8)}
        if line.startswith("CONVERGENCE") or line.startswith("ABNORMAL"):
            out["task"] = line.rstrip()
    return out


def parse_driver23(path):
    out = {"iterates": [], "final_x_str": []}
    in_x = False
    for line in open(path):
        m = re.match(r"Iterate\s+(\d+)\s+nfg =\s+(\d+)\s+f =\s*(\S+)\s+\|proj g\| =\s*(\S+)", line)
        if m:
            out["iterates"].append({"iter": int(m.group(1)), "nfg": int(m.group(2)),
                                    "f_str": m.group(3), "pg_str": m.group(4)})
            continue
        if line.strip().startswith("STOP"):
            out["task"] = line.strip()
            continue
        if "Final X=" in line:
            in_x = True
            continue
        if in_x:
            out["final_x_str"].extend(line.split())
    return out


def parse_iterate_dat(path):
    rows = []
    for line in open(path):
        p = line.split()
        if len(p) == 10 and p[0].isdigit() and p[1].isdigit():
            rows.append({"it": int(p[0]), "nf": int(p[1]), "nseg": p[2], "nact": p[3], "sub": p[4],
                         "itls": p[5], "stepl": p[6], "tstep": p[7], "projg": p[8], "f": p[9]})
    return rows


def main():
    g = {
        "source": "jacobwilliams/lbfgsb test/OUTPUTS (golden stdout of driver1/2/3, F90 and F77 builds)",
        "driver1_90": parse_driver1(os.path.join(REF, "output_90_1")),
        "driver1_77": parse_driver1(os.path.join(REF, "output_77_1")),
        "driver2_90": parse_driver23(os.path.join(REF, "output_90_2")),
        "driver2_77": parse_driver23(os.path.join(REF, "output_77_2")),
        "driver3_90": parse_driver23(os.path.join(REF, "output_90_3")),
        "driver3_77": parse_driver23(os.path.join(REF, "output_77_3")),
        "iterate_dat": parse_iterate_dat(os.path.join(REF, "iterate.dat")),
        # the raw text (lines without the trailing newline) for the end-to-end diff of the iprint output
        "driver1_90_text": [ln.rstrip("\n") for ln in open(os.path.join(REF, "output_90_1"))],
        "iterate_dat_text": [ln.rstrip("\n") for ln in open(os.path.join(REF, "iterate.dat"))],
    }
    with open(os.path.join(HERE, "reference_outputs.json"), "w") as fh:
        json.dump(g, fh, indent=1)
    print({k: (len(v["iterates"]) if isinstance(v, dict) and "iterates" in v else len(v)) for k, v in g.items() if k != "source"})


if __name__ == "__main__":
    main()
