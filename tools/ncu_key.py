#!/usr/bin/env python
"""Key metrics of every launch in an ncu report (raw page): duration, issue/pipe utilisation, memory traffic, stalls.
usage: ncu_key.py <report.ncu-rep>"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__inst_issued.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name"), "id", d.get("ID"))
    for k in KEYS:
        if k in d and d[k] != "":
            print(f"   {k:75s} {d[k]}")
    st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(v)) for k, v in d.items()
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v not in ("", "0")}
    tot = sum(st.values()) or 1
    print("   stalls:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
