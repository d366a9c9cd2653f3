#!/usr/bin/env python
"""Summaries of ncu output for profiles/:  launches <csv> | full <ncu-rep>   (run in the build container, no GPU)."""
import collections
import csv
import io
import subprocess
import sys


def launches(path):
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    agg = collections.OrderedDict()
    for r in rows:
        n = r["Kernel Name"].split("(")[0]
        n = n if len(n) < 70 else n[:67] + "..."
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"# {len(rows)} launches, {tot / 1e3:.3f} ms summed device time (cold-cache, serialised: compare shares)")
    print(f"{'kernel':70s} {'n':>5s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} {v[0]:5d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {v[1] / tot:7.3f}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[ki].split("(")[0])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:70s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
