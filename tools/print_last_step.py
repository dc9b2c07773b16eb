#!/usr/bin/env python
"""Development aid: prints the launches of the last full-path step of an ncu launch list (gpu__time_duration.sum)."""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
seq = [(r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), float(r["Metric Value"].replace(",", "")) / 1e6,
        r["Grid Size"]) for r in rows]
idx = [i for i, s in enumerate(seq) if "compose_kernel" in s[0]]
start = idx[-1] if idx else 0
end = len(seq)
for i in range(start + 1, len(seq)):
    if "pad_unfold" in seq[i][0] and i > start + 20:
        end = i
        break
tot = 0.0
for n, v, g in seq[start:end]:
    if n.startswith("native::") or n.startswith("cub::"):
        continue
    tot += v
    print(f"{n[:48]:48s} {v:9.4f} {g}")
print("total ms", tot)
