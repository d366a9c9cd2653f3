#!/usr/bin/env python
"""Developer aid: builds build_variants/<name>.so for compile-time knob sets (same sources, same ABI).
usage: build_variants.py name=-DFLAG=1,-DOTHER=2 ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "build_variants")
os.makedirs(OUT, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", os.path.join(OUT, name + ".so"),
           os.path.join(ROOT, "vanerf_b200", "csrc", "vanerf_b200.cu")] + [f for f in flags.split(",") if f]
    procs.append((name, subprocess.Popen(cmd)))
for name, p in procs:
    print(name, "->", p.wait())
