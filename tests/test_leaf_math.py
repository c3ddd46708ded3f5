"""The reference's own unit tests for the leaf arithmetic of the path (primal_closest_point.cpp:239-576,
primal_squared_distance.cpp:167-230, primal_ray_intersect.cpp:149-380, primal_boundingbox.cpp:523-569) replayed
against the oracle port, the compiled reference, and -- through the C ABI -- the device functions the kernels use."""
import numpy as np
import pytest

import kats


def _kinds(oracle):
    return ["port"] + (["reference"] if oracle.have_reference() else [])


def test_oracle_closest_point_kats(oracle):
    for kind in _kinds(oracle):
        kats.check_closest_point_tri(lambda p, t, e: oracle.closest_point_tri(p, t, e, kind))


def test_oracle_point_box_ray_box_scale_kats(oracle):
    for kind in _kinds(oracle):
        kats.check_squared_distance_point_box(lambda p, b: oracle.squared_distance_point_box(p, b, kind))
        kats.check_ray_box(lambda r, b, tol: oracle.intersect_ray_box(r, b, tol, kind))
        kats.check_box_scale(lambda b, s: oracle.box_scale(b, s, kind))


def test_oracle_port_equals_reference_on_slivers(oracle):
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built")
    q, t = kats.sliver_triangle_cloud(200_000)
    for eps in (1e-50, 1e-12):
        a, la = oracle.closest_point_tri(q, t, eps, "port")
        b, lb = oracle.closest_point_tri(q, t, eps, "reference")
        assert np.array_equal(la, lb) and np.array_equal(a, b)
    rng = np.random.default_rng(3)
    boxes = np.sort(rng.uniform(-1, 1, (50_000, 2, 3)), axis=1).reshape(-1, 6)
    pts = rng.uniform(-2, 2, (50_000, 3))
    assert np.array_equal(oracle.squared_distance_point_box(pts, boxes, "port"), oracle.squared_distance_point_box(pts, boxes, "reference"))
    rays = np.concatenate([rng.uniform(-2, 2, (50_000, 3)), rng.standard_normal((50_000, 3))], axis=1)
    rays[::7, 3] = 0.0  # axis-parallel components exercise the |n_d| <= tol branch
    rays[::11, 4] = 1e-17
    for tol in (np.finfo(np.float64).eps, 1e-9):
        assert np.array_equal(oracle.intersect_ray_box(rays, boxes, tol, "port"), oracle.intersect_ray_box(rays, boxes, tol, "reference"))
    for s in (1.000123, 1.0, 0.3, -2.0):
        assert np.array_equal(oracle.box_scale(boxes, s, "port"), oracle.box_scale(boxes, s, "reference"))


@pytest.mark.gpu
def test_gpu_leaf_math_reference_kats():
    from axom_b200 import primal
    kats.check_closest_point_tri(lambda p, t, e: primal.closest_point(p, t, e))
    kats.check_squared_distance_point_box(primal.squared_distance_point_box)
    kats.check_ray_box(lambda r, b, tol: primal.intersect_ray_box(r, b, tol))
    kats.check_box_scale(primal.box_scale)


@pytest.mark.gpu
def test_gpu_closest_point_slivers_bit_exact(oracle):
    """A16: thin / needle / collapsed triangles and on-feature query points: cp and loc equal bit for bit"""
    from axom_b200 import primal
    kind = "reference" if oracle.have_reference() else "port"
    q, t = kats.sliver_triangle_cloud(400_000, seed=11)
    for eps in (1e-50, 1e-12):
        cp, loc = primal.closest_point(q, t, eps)
        rcp, rloc = oracle.closest_point_tri(q, t, eps, kind)
        assert np.array_equal(loc, rloc)
        assert np.array_equal(cp, rcp)
    assert len(np.unique(rloc)) == 7  # every region of the state machine is reached


@pytest.mark.gpu
def test_gpu_point_box_ray_box_scale_bit_exact(oracle):
    from axom_b200 import primal
    kind = "reference" if oracle.have_reference() else "port"
    rng = np.random.default_rng(4)
    n = 200_000
    boxes = np.sort(rng.uniform(-1, 1, (n, 2, 3)), axis=1).reshape(-1, 6)
    pts = rng.uniform(-2, 2, (n, 3))
    pts[::5] = boxes[::5, :3]  # on a corner: distance exactly 0
    assert np.array_equal(primal.squared_distance_point_box(pts, boxes), oracle.squared_distance_point_box(pts, boxes, kind))
    rays = np.concatenate([rng.uniform(-2, 2, (n, 3)), rng.standard_normal((n, 3))], axis=1)
    rays[::7, 3] = 0.0
    rays[::11, 4] = 1e-17
    rays[::13, 3:] = 0.0  # zero direction: the Ray constructor falls back to (1,0,0)
    for tol in (np.finfo(np.float64).eps, 1e-9):
        assert np.array_equal(primal.intersect_ray_box(rays, boxes, tol), oracle.intersect_ray_box(rays, boxes, tol, kind))
    for s in (1.000123, 1.0, 0.3, -2.0):
        assert np.array_equal(primal.box_scale(boxes, s), oracle.box_scale(boxes, s, kind))
