"""Per-CUDA-source-line instruction / stall attribution from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python profiles/line_attr.py report.ncu-rep [units_per_launch] [topN]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = hdr = curkey = None
agg = collections.defaultdict(lambda: [0, 0, ""])
ops = collections.defaultdict(collections.Counter)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Line No":
        hdr = r
        ie, ws = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) <= max(ie, ws):
        continue
    try:
        ln = int(r[0])
        n = int(r[ie]) if r[ie] else 0
        w = int(r[ws]) if r[ws] else 0
    except ValueError:
        continue
    if r[1].strip():
        curkey = (cur, ln)
        agg[curkey][2] = r[1].strip()
        agg[curkey][0] += n
        agg[curkey][1] += w
    elif curkey is not None:
        toks = r[3].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
        ops[curkey][op.split(".")[0]] += n
tot = sum(v[0] for v in agg.values())
tw = sum(v[1] for v in agg.values())
print(f"total instructions/unit {tot / units:.1f}  stall samples {tw}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    t = " ".join(f"{o}:{c / units:.0f}" for o, c in ops[k].most_common(5))
    print(f"{k[0][:12]}:{k[1]:4d} inst/unit {v[0] / units:7.1f} stall% {100 * v[1] / max(tw, 1):4.1f}  {v[2][:64]} | {t}")
