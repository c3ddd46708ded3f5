"""GPU parity tests for the narrow phase downstream of findBoundingBoxes:
primal::intersect(Triangle3, Triangle3) and quest::findTriMeshIntersectionsBVH through the C ABI, compared
exactly (booleans, pair lists in SEQ order, degenerate ids) with the CPU oracle.  Mirrors
primal/tests/primal_triangle_intersect.cpp / quest/tests/quest_mesh_tester.cpp in structure."""
import numpy as np
import pytest

from axom_b200 import synth
import kats

pytestmark = pytest.mark.gpu


def test_tri_tri_reference_kats():
    from axom_b200 import intersect_triangles
    t1, t2, inc, want = kats.tri_tri_cases()
    for b in (False, True):
        m = inc == b
        got = intersect_triangles(t1[m], t2[m], includeBoundary=b, EPS=1e-8)
        assert np.array_equal(got, want[m]), (b, np.nonzero(got != want[m]))


@pytest.mark.parametrize("include_boundary", [False, True])
def test_tri_tri_random_pairs_bit_exact(oracle, include_boundary):
    from axom_b200 import intersect_triangles
    rng = np.random.default_rng(5)
    a = rng.random((200000, 3, 3))
    b = rng.random((200000, 3, 3)) * 0.4 + 0.3
    for eps in (1e-8, 1e-12):
        assert np.array_equal(intersect_triangles(a, b, include_boundary, eps), oracle.tri_tri_intersect(a, b, include_boundary, eps))


@pytest.mark.parametrize("include_boundary", [False, True])
def test_tri_tri_lattice_degenerate_and_coplanar(oracle, include_boundary):
    """integer-lattice triangles: shared vertices and edges, coplanar pairs, zero-area triangles"""
    from axom_b200 import intersect_triangles
    rng = np.random.default_rng(6)
    g = rng.integers(0, 3, (300000, 3, 3)).astype(np.float64)
    h = rng.integers(0, 3, (300000, 3, 3)).astype(np.float64)
    g[:100000, :, 2] = 0
    h[:100000, :, 2] = 0
    g[100000:150000, :, 0] = 1
    h[100000:150000, :, 0] = 1
    assert np.array_equal(intersect_triangles(g, h, include_boundary, 1e-8), oracle.tri_tri_intersect(g, h, include_boundary, 1e-8))
    import torch
    got = intersect_triangles(torch.from_numpy(g).cuda(), torch.from_numpy(h).cuda(), include_boundary, 1e-8)
    assert np.array_equal(got.cpu().numpy(), oracle.tri_tri_intersect(g, h, include_boundary, 1e-8))


def _two_spheres(f1, f2, shift):
    x, y, z, c = synth.icosphere(f1)
    x2, y2, z2, c2 = synth.icosphere(f2)
    X = np.concatenate([x, x2 * 0.9 + shift])
    Y = np.concatenate([y, y2 * 0.9])
    Z = np.concatenate([z, z2 * 0.9])
    C = np.concatenate([c, c2 + len(x)]).astype(np.int32)
    return X, Y, Z, C


@pytest.mark.parametrize("f1,f2", [(3, 2), (12, 9), (40, 31)])
def test_mesh_self_intersections_match_oracle(oracle, f1, f2):
    from axom_b200 import findTriMeshIntersectionsBVH
    X, Y, Z, C = _two_spheres(f1, f2, 0.3)
    C = np.concatenate([C, [[0, 0, 1], [2, 3, 2]]]).astype(np.int32)  # two degenerate cells
    want_pairs, want_deg = oracle.find_tri_mesh_intersections(X, Y, Z, C, 1e-8)
    pairs, deg = findTriMeshIntersectionsBVH(X, Y, Z, C, 1e-8)
    assert len(want_pairs) > 0
    assert np.array_equal(pairs, want_pairs)
    assert np.array_equal(deg, want_deg)


def test_mesh_tester_strategies_and_device_mesh(oracle):
    """all three find strategies (single walk + scatter, count/fill, forced overflow) give the same pair list;
    a device-resident mesh returns device-resident pairs"""
    import torch
    from axom_b200 import MeshTester
    X, Y, Z, C = _two_spheres(25, 20, 0.25)
    want_pairs, want_deg = oracle.find_tri_mesh_intersections(X, Y, Z, C, 1e-8)
    for strategy in (0, 1, 2):
        mt = MeshTester(X, Y, Z, C)
        mt.getBVH().setFindStrategy(strategy)
        assert np.array_equal(mt.findTriMeshIntersections(1e-8), want_pairs), strategy
        assert np.array_equal(mt.findTriMeshIntersections(1e-8), want_pairs), strategy  # second call re-uses buffers
        assert len(mt.degenerateIndices()) == 0
    mt = MeshTester(*(torch.from_numpy(a).cuda() for a in (X, Y, Z, C)))
    p = mt.findTriMeshIntersections(1e-8)
    assert p.is_cuda and np.array_equal(p.cpu().numpy(), want_pairs)


def test_mesh_tester_clean_and_empty_meshes(oracle):
    from axom_b200 import findTriMeshIntersectionsBVH
    x, y, z, c = synth.icosphere(10)  # watertight sphere: neighbours share vertices / edges, no intersections
    pairs, deg = findTriMeshIntersectionsBVH(x, y, z, c)
    wp, wd = oracle.find_tri_mesh_intersections(x, y, z, c)
    assert pairs.shape == (0, 2) and len(wp) == 0 and len(deg) == 0 and len(wd) == 0
    pairs, deg = findTriMeshIntersectionsBVH(x, y, z, c[:0])
    assert pairs.shape == (0, 2) and len(deg) == 0
    pairs, deg = findTriMeshIntersectionsBVH(x, y, z, c[:1])
    assert pairs.shape == (0, 2) and len(deg) == 0


def test_mesh_tester_threshold_dependence(oracle):
    """nearly touching surfaces: the fuzzy comparators make the answer depend on intersectionThreshold"""
    from axom_b200 import findTriMeshIntersectionsBVH
    X, Y, Z, C = _two_spheres(10, 10, 0.95 + 1e-9)
    for th in (1e-4, 1e-8, 1e-12):
        want, _ = oracle.find_tri_mesh_intersections(X, Y, Z, C, th)
        got, _ = findTriMeshIntersectionsBVH(X, Y, Z, C, th)
        assert np.array_equal(got, want), th
