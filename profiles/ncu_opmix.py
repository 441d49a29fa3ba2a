"""SASS opcode mix (executed warp-instructions) from an `ncu --page source --csv --print-source cuda,sass` dump.
usage: python ncu_opmix.py src.csv [N]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr, agg, tot = None, collections.Counter(), 0
seen = set()
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        ia = hdr.index("Address")
        iex = hdr.index("Instructions Executed")
        continue
    if hdr and len(r) > iex and r[ia].startswith("0x") and r[ia] not in seen:
        seen.add(r[ia])
        toks = r[ia + 1].strip().split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = ".".join(op.split(".")[:2]) if op.startswith(("SHFL", "LDS", "STS", "LDG", "STG", "BAR")) else op.split(".")[0]
        n = int(r[iex] or 0)
        agg[op] += n
        tot += n
print("executed warp-instructions:", tot)
for k, v in agg.most_common(top):
    print(f"{k:14s} {v / 1e6:10.1f} M  {100 * v / tot:5.1f}%")
