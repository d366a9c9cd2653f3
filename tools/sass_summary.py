#!/usr/bin/env python
"""Writes profiles/<round>_sass_summary.txt: per kernel of libvanerf_b200.so (sm_100a cubin), the SASS mnemonics that prove the
Blackwell-native path (B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
UBLKCP (cp.async.bulk on the TMA engine), UTMALDG / UTMASTG (tensor-map TMA), SYNCS (mbarrier), HMMA (legacy mma.sync),
FADD2 / FMUL2 / FFMA2 (packed fp32), plus instruction count and registers."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "vanerf_b200", "libvanerf_b200.so")
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = {m.group(1): m.group(2) for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", res)}
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "FADD2", "FMUL2", "FFMA2", "LDGSTS"]
cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
out = [f"# {rnd} - SASS summary of vanerf_b200/libvanerf_b200.so (sm_100a), written by tools/sass_summary.py",
       "# kernel | instructions | registers | mnemonic counts (only non-zero)"]
for fn in sorted(total, key=lambda f: -total[f]):
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
    c = " ".join(f"{k}={v}" for k, v in counts[fn].items())
    out.append(f"{name:60s} {total[fn]:6d} instr  {regs.get(fn, '?'):>3s} regs  {c}")
out.append("")
out.append("tcgen05.mma -> UTCHMMA, tcgen05.ld / st -> LDTM / STTM, tcgen05.commit -> UTCBAR, cp.async.bulk (global -> shared on the TMA engine, "
           "no tensor map: the gather writes operand images that already are the K-major 128-byte-swizzle layout, so one bulk copy per 16 KB "
           "image needs no re-layout) -> UBLKCP; there is deliberately no UTMALDG.  No HMMA: nothing runs on the legacy mma.sync path.")
p = os.path.join(ROOT, "profiles", f"{rnd}_sass_summary.txt")
open(p, "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
