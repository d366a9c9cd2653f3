"""Builds the native pieces in-tree.

  cuda  : vanerf_b200/libvanerf_b200.so  (nvcc, sm_100a only)               - the product
  oracle: oracle/_build/libgeom_oracle.so (gcc, -ffp-contract=off)           - the checker
  emul  : tests/_emul/libvanerf_emul.so  (g++ -DVANERF_HOST_EMUL)            - kernel-logic tests without a GPU
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vanerf_b200", "csrc")
CUDA_SO = os.path.join(ROOT, "vanerf_b200", "libvanerf_b200.so")
EMUL_SO = os.path.join(ROOT, "tests", "_emul", "libvanerf_emul.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "vanerf_b200.h")]


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_cuda(force=False, verbose=False):
    if not force and not _stale(CUDA_SO):
        return CUDA_SO
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", CUDA_SO, os.path.join(CSRC, "vanerf_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return CUDA_SO


def build_trace(force=True):
    """Developer build with the in-kernel cycle trace of k_mlp_tc compiled in (tools/tc_trace.py); not the product."""
    out = os.path.join(ROOT, "vanerf_b200", "libvanerf_b200_trace.so")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DVANERF_TC_TRACE",
                           "-Xcompiler", "-fPIC", "-shared", "-o", out, os.path.join(CSRC, "vanerf_b200.cu")])
    return out


def build_emul(force=False):
    if not force and not _stale(EMUL_SO):
        return EMUL_SO
    os.makedirs(os.path.dirname(EMUL_SO), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-DVANERF_HOST_EMUL", "-x", "c++", "-shared",
                           "-fPIC", "-pthread", "-o", EMUL_SO, os.path.join(CSRC, "vanerf_b200.cu"),
                           os.path.join(CSRC, "host_emul.cpp")])
    return EMUL_SO


def build_oracle(force=False):
    sys.path.insert(0, ROOT)
    from oracle import geom
    return geom.build(force)


if __name__ == "__main__":
    what = sys.argv[1:] or ["cuda", "oracle", "emul"]
    for w in what:
        print(w, "->", {"cuda": build_cuda, "oracle": build_oracle, "emul": build_emul, "trace": build_trace}[w](force=True))
