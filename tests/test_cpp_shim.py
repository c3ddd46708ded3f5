"""The C++ header shims (include/axom_b200/*.hpp, traverser.cuh) compiled and run as a user would:
plain g++ / nvcc against libaxb200.so.  CPU: they compile and the no-device path fails loudly.
GPU: the reference's spin_bvh.cpp / signed_distance KATs through the shim, and a third-party kernel
walking getTraverser()'s arrays reproduces findPoints exactly."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "bin")


def _build():
    import __graft_entry__ as g
    g.build_cpp_tests()
    return os.path.join(BIN, "shim_test"), os.path.join(BIN, "traverser_test")


def test_shims_compile_and_fail_loudly_without_device():
    shim, trav = _build()
    assert os.path.exists(shim) and os.path.exists(trav)
    r = subprocess.run([shim, "--no-device"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_shim_kats_on_gpu():
    shim, _ = _build()
    r = subprocess.run([shim], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "shim_test: OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_user_kernel_with_traverser_on_gpu():
    _, trav = _build()
    r = subprocess.run([trav], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "traverser_test: OK" in r.stdout, r.stdout + r.stderr
