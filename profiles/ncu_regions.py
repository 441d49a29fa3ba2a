"""Stall samples of an `ncu --page source --csv` dump grouped into code REGIONS by execution count (the inner loops of a
kernel execute far more often than its per-slice code), with the FP64 instruction count of each region.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python ncu_regions.py src.csv [N regions]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 5
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ix["Instructions Executed"]]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
print(rows[0][1], "-- total stall samples", tot)
groups = collections.defaultdict(list)
for r in data:
    groups[int(r[ix["Instructions Executed"]] or 0)].append(r)
for ex, rs in sorted(groups.items(), key=lambda kv: -sum(int(r[ix["# Samples"]] or 0) for r in kv[1]))[:top]:
    smp = sum(int(r[ix["# Samples"]] or 0) for r in rs)
    agg, ops = collections.Counter(), collections.Counter()
    for r in rs:
        for s in stalls:
            agg[s.replace("stall_", "")] += int(r[ix[s]] or 0)
        toks = r[ix["Source"]].split()
        ops[(toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]] += 1
    fp64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
    issue = agg["selected"] / smp if smp else 0
    print(f"region executed {ex} x per warp-instruction: {len(rs)} SASS instructions ({fp64} FP64), "
          f"{100 * smp / tot:.1f} % of the samples, issuing {100 * issue:.0f} % of its cycles"
          + (f" -> {len(rs) / issue / fp64:.2f} cycles per FP64 instruction" if fp64 and issue else ""))
    print("    stalls:", ", ".join(f"{k} {100 * v / smp:.0f} %" for k, v in agg.most_common(5)))
    print("    opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(8)))
