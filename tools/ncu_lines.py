#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples from an ncu report (needs -lineinfo + --import-source on), with
the dominant stall reasons of each line, plus the kernel-wide stall mix.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> [top N]"""
import collections
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
stall_tot = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
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
    e = lines.setdefault(key, [0, 0, r[1], collections.Counter()])
    e[0] += smp
    e[1] += ex
    for i, name in stall_cols:
        try:
            v = int(r[i])
        except ValueError:
            continue
        e[3][name] += v
        stall_tot[name] += v
tot = sum(v[0] for v in lines.values())
print(f"# {kern}: {tot} samples over {len(lines)} source lines (summed over the captured launches); "
      f"{sum(v[1] for v in lines.values())} warp instructions")
print("# stall mix:", ", ".join(f"{k} {100.0 * v / max(sum(stall_tot.values()), 1):.1f}%" for k, v in stall_tot.most_common(10)))
for (f, ln), (smp, ex, src, st) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    why = " ".join(f"{k}:{v}" for k, v in st.most_common(3))
    print(f"{smp:7d} {100.0 * smp / max(tot, 1):5.1f}%  inst {ex:9d}  {f}:{ln}: {src.strip()[:70]}   [{why}]")
