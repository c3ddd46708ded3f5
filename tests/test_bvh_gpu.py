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


def test_build_1m(oracle):
    boxes = synth.triangle_aabbs(1_000_000, seed=12345)
    _check_build(oracle, boxes, 3)


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
