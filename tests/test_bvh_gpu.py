"""GPU parity tests for spin::BVH: every call goes through the C ABI (libaxb200.so) and is
compared bit for bit with the CPU oracle on the same seeded inputs.  Mirrors the structure of
the reference's spin/tests/spin_bvh.cpp (build bounds, box / ray / point queries, N=0 and N=1)."""
import numpy as np
import pytest

from axom_b200 import synth

pytestmark = pytest.mark.gpu


def _gpu_bvh(boxes, ndims, scale=None, tol=None):
    from axom_b200 import BVH
    b = BVH(ndims)
    if scale is not None:
        b.setScaleFactor(scale)
    if tol is not None:
        b.setTolerance(tol)
    assert b.initialize(boxes) == 0
    return b


def _check_build(oracle, boxes, ndims, scale=None):
    ref = oracle.Bvh(boxes, ndims=ndims, scale=-1.0 if scale is None else scale)
    gpu = _gpu_bvh(boxes, ndims, scale)
    A, G = ref.arrays(), gpu.arrays()
    assert gpu.numLeaves() == ref.n
    for k in ("mcodes", "leafs", "inner_children", "inner_nodes", "bounds"):
        assert np.array_equal(A[k], G[k]), (k, ndims, len(boxes))
    return ref, gpu


@pytest.mark.parametrize("ndims", [3, 2])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 27, 1000, 4097, 50000])
def test_build_bit_exact(oracle, ndims, n):
    boxes = synth.triangle_aabbs(max(n, 1), seed=100 + n, ndims=ndims)[:n]
    if n > 10:
        boxes[5] = boxes[6]  # duplicate box -> tied Morton codes, resolved by index
    _check_build(oracle, boxes, ndims)
    _check_build(oracle, boxes, ndims, scale=1.0)


def test_build_many_ties(oracle):
    # clustered boxes: thousands of identical Morton codes (10 bits/dim)
    rng = np.random.default_rng(3)
    c = rng.random((20000, 3)) * 1e-4 + 0.5
    boxes = np.concatenate([c - 1e-6, c + 1e-6], axis=1)
    boxes[0, :3] = 0.0
    boxes[1, 3:] = 1.0
    _check_build(oracle, boxes, 3)


def test_build_legacy_path(oracle, monkeypatch):
    """AXB_BUILD_LEGACY=1: the reference-order top-down tree_kernel + refit_kernel pair (the fallback of the
    fused bottom-up build) must give the same tree."""
    monkeypatch.setenv("AXB_BUILD_LEGACY", "1")
    for n, nd in ((4097, 3), (50000, 2), (200000, 3)):
        boxes = synth.triangle_aabbs(n, seed=31 + n, ndims=nd)
        boxes[100:140] = boxes[100]
        _check_build(oracle, boxes, nd)


@pytest.mark.parametrize("block", [128, 512])
def test_build_fused_block_sizes(oracle, monkeypatch, block):
    """the fused build with other leaves-per-block settings (default 256): in-block vs cross-block merges move"""
    monkeypatch.setenv("AXB_AGGLO_BLOCK", str(block))
    boxes = synth.triangle_aabbs(70001, seed=block)
    boxes[1000:1300] = boxes[1000]
    _check_build(oracle, boxes, 3)
    _check_build(oracle, synth.triangle_aabbs(33333, seed=block + 1, ndims=2), 2)


def test_build_noncanonical_invalid_boxes(oracle):
    """input boxes with min > max in one dimension only (not the canonical invalid box): the fused build hands
    over to the reference-order kernels, whose addBox handling is the reference's (BoundingBox.hpp:487-508)"""
    boxes = synth.triangle_aabbs(5000, seed=11)
    boxes[17, 0], boxes[17, 3] = 0.9, 0.1
    boxes[4000, 1], boxes[4000, 4] = 0.8, 0.2
    _check_build(oracle, boxes, 3)


def test_build_sorted_and_reversed_input(oracle):
    """inputs already in Morton order / reverse order: long runs of in-block merges and deep cross-block chains"""
    boxes = synth.triangle_aabbs(300000, seed=5)
    ref = oracle.Bvh(boxes, ndims=3)
    order = ref.arrays()["leafs"]
    _check_build(oracle, np.ascontiguousarray(boxes[order]), 3)
    _check_build(oracle, np.ascontiguousarray(boxes[order[::-1]]), 3)


def test_build_1m(oracle):
    boxes = synth.triangle_aabbs(1_000_000, seed=12345)
    _check_build(oracle, boxes, 3)


def test_build_c4_20m_surface(oracle, have_ref):
    """BASELINE config 4's surface (20 M leaves): the only size at which build_radix_tree.hpp:336-339's float32(l)
    rounding (above 2^24 leaves) and the 30-deep Morton tie chains of a sphere quantised to 10 bits per axis are
    reachable.  Compared with the UNMODIFIED reference (oracle/_ref) when it is present, else with the port."""
    x, y, z, conn = synth.icosphere(1000)
    P = np.stack([x, y, z], axis=1)
    t = P[conn]
    boxes = np.ascontiguousarray(np.concatenate([t.min(axis=1), t.max(axis=1)], axis=1))
    del t
    assert len(boxes) == 20_000_000 > (1 << 24)
    ref = oracle.Bvh(boxes, ndims=3, kind="reference" if have_ref else "port")
    gpu = _gpu_bvh(boxes, 3)
    A, G = ref.arrays(), gpu.arrays()
    # heavy ties: this is what makes the case interesting
    assert np.unique(A["mcodes"]).size < 0.5 * len(boxes)
    for k in ("mcodes", "leafs", "inner_children", "bounds", "inner_nodes"):
        assert np.array_equal(A[k], G[k]), k


def test_soa_and_device_inputs(oracle):
    import torch
    boxes = synth.triangle_aabbs(5000, seed=9)
    ref = oracle.Bvh(boxes, ndims=3).arrays()
    from axom_b200 import BVH
    soa = tuple(np.ascontiguousarray(boxes[:, c]) for c in range(6))  # ZipIndexable form
    for inp in (soa, torch.from_numpy(boxes).cuda(), tuple(torch.from_numpy(a).cuda() for a in soa)):
        b = BVH(3)
        assert b.initialize(inp) == 0
        G = b.arrays()
        for k in ("mcodes", "leafs", "inner_children", "inner_nodes", "bounds"):
            assert np.array_equal(ref[k], G[k]), k


def _same(a, b):
    return all(np.array_equal(np.asarray(u), np.asarray(v)) for u, v in zip(a, b))


@pytest.mark.parametrize("ndims", [3, 2])
@pytest.mark.parametrize("n", [0, 1, 2, 27, 20000])
def test_find_queries_bit_exact(oracle, ndims, n):
    boxes = synth.triangle_aabbs(max(n, 1), seed=7 + n, ndims=ndims)[:n]
    ref, gpu = _check_build(oracle, boxes, ndims)
    pts = synth.random_points(3000, seed=n, ndims=ndims)
    assert _same(ref.find_points(pts), gpu.findPoints(pts))
    qb = synth.triangle_aabbs(2500, seed=n + 99, ndims=ndims)
    assert _same(ref.find_boxes(qb), gpu.findBoundingBoxes(qb))
    o, d = synth.random_rays(1200, seed=n + 3, lo=-0.5, hi=1.5, ndims=ndims)
    assert _same(ref.find_rays(o, d * 1.7, True), gpu.findRays(o, d * 1.7, normalized=False))
    assert _same(ref.find_rays(o, d, False), gpu.findRays(o, d, normalized=True))


@pytest.mark.parametrize("strategy", [0, 1, 2])
def test_find_strategies_large_batches(oracle, strategy):
    """enough queries for the Morton-sorted path (>= 8192); strategy 0 = single traversal + scatter,
    1 = the reference's count/fill double traversal, 2 = forced overflow of the hit buffer (fallback).
    All three must reproduce the reference's offsets / counts / candidates exactly, order included."""
    boxes = synth.triangle_aabbs(60000, seed=77)
    ref, gpu = _check_build(oracle, boxes, 3)
    gpu.setFindStrategy(strategy)
    pts = synth.random_points(50000, seed=3)
    pts[:100] += 5.0  # outside the tree bounds
    assert _same(ref.find_points(pts), gpu.findPoints(pts))
    qb = synth.triangle_aabbs(20000, seed=78)
    qb[:, 3:] += 0.05  # fat query boxes: tens of candidates each
    assert _same(ref.find_boxes(qb), gpu.findBoundingBoxes(qb))
    o, d = synth.random_rays(12000, seed=79, lo=-0.2, hi=1.2)
    assert _same(ref.find_rays(o, d, True), gpu.findRays(o, d))
    # second call re-uses the sized buffers
    assert _same(ref.find_boxes(qb), gpu.findBoundingBoxes(qb))


@pytest.mark.parametrize("compact", ["1", "0"])
def test_find_compact_records_boundary_cases(oracle, monkeypatch, compact):
    """point / box walks use 64-byte records with outward-rounded float boxes (AXB_FIND_COMPACT=0: the 128-byte ones);
    leaves are confirmed on the exact double boxes, so queries that touch a box exactly, miss it by one ulp, or fall in
    the float rounding gap give the reference's candidates in both settings"""
    monkeypatch.setenv("AXB_FIND_COMPACT", compact)
    rng = np.random.default_rng(12)
    n = 40000
    c = rng.random((n, 3)) * 1000.0 + 1.0 / 3.0  # coordinates that are not representable in binary32
    boxes = np.concatenate([c, c + rng.random((n, 3)) * 3.0], axis=1)
    ref, gpu = _check_build(oracle, boxes, 3, scale=1.0)
    # points on box corners, one ulp inside / outside, and in the float rounding gap around them
    corner = np.concatenate([boxes[:6000, :3], boxes[6000:12000, 3:]])
    pts = np.concatenate([corner, np.nextafter(corner, np.inf), np.nextafter(corner, -np.inf),
                          corner.astype(np.float32).astype(np.float64), corner * (1 + 1e-9)])
    assert _same(ref.find_points(pts), gpu.findPoints(pts))
    # query boxes that touch stored boxes exactly or miss them by one ulp
    lo, hi = boxes[:9000, 3:], boxes[:9000, 3:] + 2.0
    qb = np.concatenate([np.concatenate([lo, hi], 1), np.concatenate([np.nextafter(lo, np.inf), hi], 1),
                         np.concatenate([boxes[9000:12000, :3] - 2.0, np.nextafter(boxes[9000:12000, :3], -np.inf)], 1)])
    assert _same(ref.find_boxes(qb), gpu.findBoundingBoxes(qb))
    # default scale factor (inflated leaf boxes) as well
    ref2, gpu2 = _check_build(oracle, boxes, 3)
    assert _same(ref2.find_points(pts), gpu2.findPoints(pts))
    assert _same(ref2.find_boxes(qb), gpu2.findBoundingBoxes(qb))


def test_find_device_resident(oracle):
    import torch
    boxes = synth.triangle_aabbs(30000, seed=21)
    ref, gpu = _check_build(oracle, boxes, 3)
    pts = synth.random_points(10000, seed=5)
    off, cnt, cand = gpu.findPoints(torch.from_numpy(pts).cuda())
    r = ref.find_points(pts)
    assert _same(r, (off.cpu().numpy(), cnt.cpu().numpy(), cand.cpu().numpy()))


def test_query_before_build_fails():
    from axom_b200 import BVH
    from axom_b200._lib import AxbError, AXB_ERR_NOT_BUILT
    b = BVH(3)
    assert not b.isInitialized()
    lo, hi = b.getBounds()
    assert (lo > hi).all()  # invalid box, spin/BVH.hpp:303-307
    with pytest.raises(AxbError) as e:
        b.findPoints(np.zeros((1, 3)))
    assert e.value.status == AXB_ERR_NOT_BUILT


class _Adapter:
    """gives axom_b200.BVH the oracle.Bvh query interface so tests/kats.py runs on the GPU path"""

    def __init__(self, boxes, ndims, scale):
        self.b = _gpu_bvh(boxes, ndims, scale)

    def find_points(self, p):
        return self.b.findPoints(p)

    def find_boxes(self, q):
        return self.b.findBoundingBoxes(q)

    def find_rays(self, o, d, normalize=True):
        return self.b.findRays(o, d, normalized=not normalize)


def test_reference_kats_on_gpu():
    import kats
    b3 = kats.unit_cells(3, 3)
    g = _Adapter(b3, 3, 1.0)
    A = g.b.arrays()
    assert A["mcodes"].tolist() == kats.KAT_MCODES and A["leafs"].tolist() == kats.KAT_LEAFS
    l, r = kats.children_to_lr(A["inner_children"], 27)
    assert l.tolist() == kats.KAT_LCHILD and r.tolist() == kats.KAT_RCHILD
    kats.check_boxes_3d(g)
    kats.check_rays_3d(g)
    kats.check_points(g, b3, 3)
    b2 = kats.unit_cells(3, 2)
    g2 = _Adapter(b2, 2, 1.0)
    kats.check_boxes_2d(g2)
    kats.check_rays_2d(g2)
    kats.check_points(g2, b2, 2)
    lo, hi = _gpu_bvh(b3, 3).getBounds()  # default scale 1.000123 inflates the bounds
    assert lo[0] == -6.1500000000047628e-05 and hi[0] == 3.0000615000000002


@pytest.mark.parametrize("name", ["bvh3d_n600", "bvh2d_n400"])
def test_golden_fixtures_on_gpu(name):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    nd = int(g["ndims"])
    b = _gpu_bvh(g["boxes"], nd)
    A = b.arrays()
    for k in ("mcodes", "leafs", "inner_children", "inner_nodes", "bounds"):
        assert np.array_equal(A[k], g["a_" + k]), k
    assert _same(b.findPoints(g["pts"]), (g["p_off"], g["p_cnt"], g["p_cand"]))
    assert _same(b.findBoundingBoxes(g["qboxes"]), (g["b_off"], g["b_cnt"], g["b_cand"]))
    assert _same(b.findRays(g["ray_o"], g["ray_d"]), (g["r_off"], g["r_cnt"], g["r_cand"]))


def test_rebuild_and_traverser_view():
    import torch
    from axom_b200 import BVH
    b = BVH(3)
    b.initialize(synth.triangle_aabbs(100, seed=1))
    boxes = synth.triangle_aabbs(7000, seed=2)
    b.initialize(boxes)  # initialize() may be called again (spin/BVH.hpp:437)
    t = b.getTraverser()
    assert t.num_leaves == 7000 and t.ndims == 3 and t.fp_bytes == 8
    assert t.inner_nodes and t.inner_node_children and t.leaf_nodes


@pytest.mark.parametrize("ndims", [3, 2])
@pytest.mark.parametrize("n", [0, 1, 2, 27, 5000, 60000])
def test_float_bvh_bit_exact(oracle, have_ref, ndims, n):
    """FloatType = float (spin/tests/spin_bvh.cpp:1563-1669 instantiates it): build artefacts and all three
    queries against the float oracle -- the real reference's BVH<D,SEQ_EXEC,float> when oracle/_ref is
    present, and always the float build of the restatement."""
    from axom_b200 import BVH
    boxes = synth.triangle_aabbs(max(n, 1), seed=5 + n, ndims=ndims)[:n].astype(np.float32)
    pts = synth.random_points(9000, seed=n, ndims=ndims).astype(np.float32)
    qb = synth.triangle_aabbs(1500, seed=n + 9, ndims=ndims).astype(np.float32)
    o, d = synth.random_rays(800, seed=n + 3, lo=-0.5, hi=1.5, ndims=ndims)
    o, d = o.astype(np.float32), d.astype(np.float32)
    for scale in (None, 1.0):
        gpu = BVH(ndims, dtype=np.float32)
        if scale is not None:
            gpu.setScaleFactor(scale)
        assert gpu.getTolerance() == float(np.finfo(np.float32).eps)
        assert gpu.initialize(boxes) == 0
        G = gpu.arrays()
        for kind in (["reference_f32"] if have_ref else []) + ["port_f32"]:
            ref = oracle.Bvh(boxes, ndims, scale=-1.0 if scale is None else scale, kind=kind)
            A = ref.arrays()
            for k in ("mcodes", "leafs", "inner_children", "inner_nodes", "bounds"):
                assert np.array_equal(A[k], G[k]), (k, kind, ndims, n)
            assert _same(ref.find_points(pts), gpu.findPoints(pts))
            assert _same(ref.find_boxes(qb), gpu.findBoundingBoxes(qb))
            assert _same(ref.find_rays(o, d * np.float32(1.7), True), gpu.findRays(o, d * np.float32(1.7), normalized=False))
            assert _same(ref.find_rays(o, d, False), gpu.findRays(o, d, normalized=True))


@pytest.mark.parametrize("nd", [3, 2])
def test_write_vtk_file_matches_the_reference_byte_for_byte(oracle, have_ref, tmp_path, nd):
    """BVH::writeVtkFile (spin/BVH.hpp:405, policy/LinearBVH.hpp:404-458): same boxes, same order, same stream formatting as the
    unmodified reference (golden: tests/golden/bvh_vtk_*.vtk, made by tests/golden/make_golden.py; compared with the
    reference itself when it is present), double and float trees."""
    import os
    import kats
    from axom_b200 import BVH
    boxes = kats.vtk_boxes(nd)
    b = BVH(nd, device=0)
    assert b.initialize(boxes) == 0
    f = str(tmp_path / "gpu.vtk")
    b.writeVtkFile(f)
    want = open(os.path.join(os.path.dirname(__file__), "golden", "bvh_vtk_%dd.vtk" % nd)).read()
    assert open(f).read() == want
    if have_ref:
        for kind, dt in (("reference", np.float64), ("reference_f32", np.float32)):
            r = str(tmp_path / (kind + ".vtk"))
            oracle.Bvh(boxes, ndims=nd, kind=kind).write_vtk(r)
            g = BVH(nd, device=0, dtype=dt)
            assert g.initialize(boxes.astype(dt)) == 0
            g.writeVtkFile(f)
            assert open(f).read() == open(r).read(), kind
