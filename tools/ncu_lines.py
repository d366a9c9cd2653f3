#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples from an ncu report (needs -lineinfo + --import-source on).
usage: ncu_lines.py <report.ncu-rep> <kernel regex> [top N]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, {}
first_launch_done = False
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        ln = int(r[0])
        smp = int(r[hdr.index("# Samples")])
        ex = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    key = (cur_file, ln)
    e = lines.setdefault(key, [0, 0, r[1]])
    e[0] += smp
    e[1] += ex
tot = sum(v[0] for v in lines.values())
print(f"# {kern}: {tot} samples over {len(lines)} source lines (summed over the captured launches)")
for (f, ln), (smp, ex, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{smp:7d} {100.0 * smp / max(tot, 1):5.1f}%  inst {ex:10d}  {f}:{ln}: {src.strip()[:110]}")
