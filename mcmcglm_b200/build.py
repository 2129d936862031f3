"""Builds libcggibbs.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(CSRC, "libcggibbs.so")
SOURCES = ["cggibbs.cu"]
DEPS = ["cggibbs.cu", "cgg_device.cuh", "cgg_math.cuh", "cgg_jet.cuh", "cgg_math_tables.cuh", os.path.join("..", "..", "include", "cggibbs.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-ldl"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    extra = os.environ.get("CGG_NVCC_EXTRA", "").split()      # experiments: e.g. -DCGG_THREADS=640
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libcggibbs.so")
    with open(os.path.join(CSRC, "ptxas_info.txt"), "w") as f:
        f.write(r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose=True))
