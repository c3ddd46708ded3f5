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


@pytest.mark.gpu
def test_point_overload_enters_nearer_centroid_first_on_gpu(oracle, have_ref, tmp_path):
    """getTraverser().traverse_tree(Point, ...) (spin/policy/LinearBVH.hpp:72-85) in a user kernel, with the
    nearest-neighbour lambdas and the 101 x 101 query lattice of spin_bvh.cpp:1401-1554.  Source points on a lattice,
    so that up to four are equidistant from a query: which one a strict-< leaf action keeps depends on the centroid
    ordering.  Compared with the unmodified reference's own traverser (oracle/_ref) when present, else with the port
    (pinned to it by tests/test_distributed_closest_point.py)."""
    import numpy as np
    _, trav = _build()
    g = np.linspace(-1.5, 1.5, 101)
    X, Y = np.meshgrid(g, g, indexing="xy")
    query = np.ascontiguousarray(np.stack([X.ravel(), Y.ravel()], axis=1))  # make_query_points_2d: j outer, i inner
    rng = np.random.default_rng(7)
    h = np.linspace(-1.2, 1.2, 41)
    lattice = np.stack(np.meshgrid(h, h, indexing="ij"), axis=-1).reshape(-1, 2)
    cases = {"one point (the reference's own case)": np.array([[0.45, 0.8]]),
             "lattice with ties": lattice,
             "shuffled lattice": lattice[rng.permutation(len(lattice))],
             "random cloud": rng.uniform(-1.5, 1.5, (3000, 2)),
             "no points": np.empty((0, 2))}
    kind = "reference" if have_ref else "port"
    ties = 0
    for name, src in cases.items():
        src = np.ascontiguousarray(src, np.float64)
        fs, fq, fo = (str(tmp_path / f) for f in ("src.bin", "q.bin", "out.bin"))
        src.tofile(fs)
        query.tofile(fq)
        r = subprocess.run([trav, "--nearest2d", fs, fq, fo], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, name + ": " + r.stdout + r.stderr
        raw = np.fromfile(fo, np.uint8)
        nq = len(query)
        elem = raw[:4 * nq].view(np.int32)
        sq = raw[4 * nq:].view(np.float64)
        if len(src) == 0:
            assert (elem == -1).all()
            continue
        want = oracle.DistributedClosestPointRank(src, ndims=2, kind=kind).compute_local(0, query)
        assert np.array_equal(elem, want["cp_index"]), name
        assert np.array_equal(np.sqrt(sq), want["cp_distance"]), name
        d2 = ((query[:, None, :] - src[None, :, :]) ** 2).sum(-1) if len(src) <= 3000 else None
        if d2 is not None:
            ties += int((np.isclose(d2, d2.min(axis=1, keepdims=True), rtol=0, atol=0).sum(axis=1) > 1).sum())
    assert ties > 100  # the lattice cases really do exercise the tie-break
