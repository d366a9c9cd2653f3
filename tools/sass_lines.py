#!/usr/bin/env python
"""Static SASS statistics of one kernel per source line: local-memory ops (spills / stack) and instruction counts.
usage: sass_lines.py <kernel substring> [top N]   (compiles vanerf_b200.cu to a cubin under /tmp first)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kern = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cubin = "/tmp/vanerf_sass.cubin"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin", "-o", cubin,
                       os.path.join(ROOT, "vanerf_b200", "csrc", "vanerf_b200.cu")])
out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
cur, inside = None, False
loc, tot, ops = collections.Counter(), collections.Counter(), collections.Counter()
for l in out:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        inside = kern in m.group(1)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        tot[cur] += 1
        ops[m.group(1).split(".")[0]] += 1
        if re.match(r"(LDL|STL)", m.group(1)):
            loc[cur] += 1
print(f"{kern}: {sum(tot.values())} instructions, {sum(loc.values())} local-memory ops")
print("opcode mix:", ", ".join(f"{k} {v}" for k, v in ops.most_common(30)))
print("--- local-memory ops by line")
for k, v in loc.most_common(top):
    print("  ", k, v)
print("--- instructions by line")
for k, v in tot.most_common(top):
    print("  ", k, v)
