"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python ncu_lines.py src.csv [N]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
fname, line = "", None
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; ia = hdr.index("Address"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); continue
    if hdr is None:
        continue
    if r[0] != "":
        line = (fname, int(r[0])); agg[line][2] = r[1]; continue
    if len(r) > iex and r[ia].startswith("0x") and line is not None:
        agg[line][0] += int(r[isamp] or 0); agg[line][1] += int(r[iex] or 0)
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
print("samples", tot, "warp-instructions", toti)
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot:5.1f}% smp {100*v[1]/toti:5.1f}% inst {ln[0]}:{ln[1]:<4d} {v[2].strip()[:95]}")
