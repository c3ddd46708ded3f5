#!/usr/bin/env python3
"""Tuning builds of the library with extra -D flags: axom_b200/lib/var_<name>.so (git-ignored; they travel with gpurun).
   python tools/build_variant.py b5 -DAXB_SD2_MIN_BLOCKS=5 ...     then   AXB200_LIB=axom_b200/lib/var_b5.so python tools/sd_probe.py"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from axom_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out = os.path.join(B.LIBDIR, "var_%s.so" % name)
    cmd = ["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + flags + ["-ccbin", "/usr/bin/g++", "-o", out] + [s for s in B.sources() if s.endswith(".cu")]
    subprocess.check_call(cmd)
    print(out)


if __name__ == "__main__":
    main()
