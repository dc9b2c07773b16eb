#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python profiles/summarize_ncu.py launches <launches.csv> <out.txt> [skip_until_kernel_substring]
    python profiles/summarize_ncu.py full <report.ncu-rep> <out.txt>
"""
import collections
import csv
import subprocess
import sys


def launches(path, out, tail_from=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    seq = []
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6}.get(r["Metric Unit"], 1.0)
        seq.append((r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), v))
    agg = collections.OrderedDict()
    for n, v in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in seq)
    with open(out, "w") as f:
        f.write(f"# {path}: {len(seq)} launches, {tot:.3f} ms total (ncu per-launch times: cold-cache, serialised)\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg ms':>9s}\n")
        for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{n[:60]:60s} {c:8d} {v:10.3f} {100 * v / tot:6.1f}% {v / c:9.4f}\n")
        f.write("\n# last 40 launches in order\n")
        for n, v in seq[-40:]:
            f.write(f"{n[:70]:70s} {v:9.4f} ms\n")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    with open(out, "w") as f:
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write(f"## {d.get('Kernel Name', '?')}\n")
            for i, h in enumerate(hdr):
                if h in KEYS or 'pipe_tensor_cycles_active' in h or 'subpipe_hmma_cycles_active' in h:
                    f.write(f"{h:80s} {vals[i]:>18s} {units[i]}\n")
            f.write("\n")
        srows = list(csv.reader(src.splitlines()))
        if len(srows) > 2:
            h2 = srows[1]
            idx = {h: i for i, h in enumerate(h2)}
            tot = collections.Counter()
            for r in srows[2:]:
                if len(r) < len(h2):
                    continue
                for h in h2:
                    if h.startswith("stall_") and "Not Issued" not in h:
                        try:
                            tot[h] += float(r[idx[h]])
                        except ValueError:
                            pass
            s = sum(tot.values()) or 1
            f.write("## warp stall sampling (all samples, first kernel)\n")
            for k, v in tot.most_common(10):
                f.write(f"{k:30s} {100 * v / s:5.1f}%\n")
            samp = []
            for r in srows[2:]:
                if len(r) >= len(h2):
                    try:
                        samp.append((float(r[idx["# Samples"]] or 0), r[idx["Source"]].strip(), r[idx["Instructions Executed"]]))
                    except ValueError:
                        pass
            samp.sort(reverse=True)
            ts = sum(x[0] for x in samp) or 1
            f.write("\n## hottest SASS instructions by samples\n")
            for x in samp[:25]:
                f.write(f"{100 * x[0] / ts:5.1f}%  exec={x[2]:>12s}  {x[1][:100]}\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3])
