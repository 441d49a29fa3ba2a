"""Writes profiles/r2_ncu_traffic.json (read by bench.py for `roofline.traffic`) from `ncu --page raw --csv` dumps:
    python profiles/ncu_traffic.py <kernel name as bench.py prints it> <samples> <threads> <dtype> raw.csv [source note]
DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum of that one launch."""
import csv
import json
import os
import sys

kernel, samples, threads, dtype, path = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
note = sys.argv[6] if len(sys.argv) > 6 else path
rows = list(csv.reader(open(path)))
hdr, units, vals = rows[0], rows[1], rows[-1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
got = {}
for h, u, v in zip(hdr, units, vals):
    if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        got[h] = float(v.replace(",", "")) * scale[u]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2_ncu_traffic.json")
rec = json.load(open(out)) if os.path.exists(out) else {"captures": []}
rec["captures"] = [c for c in rec["captures"] if not (c["kernel"] == kernel and c["dtype"] == dtype)]
rec["captures"].append({"kernel": kernel, "dtype": dtype, "samples": samples, "threads": threads,
                        "dram_bytes_read": got["dram__bytes_read.sum"], "dram_bytes_write": got["dram__bytes_write.sum"],
                        "source": note})
json.dump(rec, open(out, "w"), indent=1)
print(open(out).read())
