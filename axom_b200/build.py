"""Compile the sm_100a CUDA library in-tree: axom_b200/lib/libaxb200.so.

    python -m axom_b200.build [--force]

nvcc cross-compiles without a GPU.  -fmad=false keeps every mul/add separately rounded
(parity with the reference's x86-64 build); -lineinfo lets ncu map SASS back to source.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libaxb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-fmad=false", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(os.path.join(SRC, f) for f in os.listdir(SRC))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(HERE, "..", "include", h) for h in ("axb200.h", "axb200_quest.h")]
    return any(os.path.getmtime(s) > t for s in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):  # e.g. a box without the toolkit: keep the shipped binary
            return LIB
        raise RuntimeError("nvcc not found at %s and no prebuilt %s" % (nvcc, LIB))
    os.makedirs(LIBDIR, exist_ok=True)
    extra = os.environ.get("AXB_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-ccbin", os.environ.get("AXB_HOST_CXX", "/usr/bin/g++"), "-o", LIB] + [s for s in sources() if s.endswith(".cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
