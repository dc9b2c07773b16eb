#!/usr/bin/env python
"""profiles/r02_traffic.json from an ncu CSV of `bench.py` (metrics dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum; kernel filter = the dominant kernel): DRAM bytes per launch of the dominant kernel over the
launches of ONE bench step (the last `per_step` launches of the capture), which bench.py reports as roofline.traffic.

    python tools/ncu_traffic.py <ncu.csv> <workload:op key> <launches per step> [out.json]
"""
import csv
import json
import os
import sys


def main():
    path, key, per_step = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
    lines = [l for l in open(path) if not l.startswith("==")]
    per_launch = {}
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        else:
            v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(u, 1)
        per_launch.setdefault(int(r["ID"]), {})[r["Metric Name"]] = v
    ids = sorted(per_launch)[-per_step:]
    rd = sum(per_launch[i].get("dram__bytes_read.sum", 0.0) for i in ids)
    wr = sum(per_launch[i].get("dram__bytes_write.sum", 0.0) for i in ids)
    ms = sum(per_launch[i].get("gpu__time_duration.sum", 0.0) for i in ids)
    table = {}
    if os.path.exists(out):
        table = json.load(open(out))
    table[key] = (rd + wr) / len(ids)
    table[key + ":detail"] = {"launches": len(ids), "dram_read_bytes": rd, "dram_write_bytes": wr, "ncu_ms_total": ms,
                              "source": os.path.basename(path)}
    json.dump(table, open(out, "w"), indent=1, sort_keys=True)
    print(key, "traffic per launch", table[key], table[key + ":detail"])


if __name__ == "__main__":
    main()
