"""Host-side mirror of the primal operators on the hot path, over the C ABI (n independent items per call):
    closest_point(Point, Triangle, loc, EPS)   primal/operators/closest_point.hpp:162-290
    squared_distance(Point, BoundingBox)       primal/operators/squared_distance.hpp:77-100
    detail::intersect_ray(Ray, BoundingBox)    primal/operators/detail/intersect_ray_impl.hpp:321-351 (the findRays predicate)
    BoundingBox::scale                         primal/geometry/BoundingBox.hpp:548-561
They run the DEVICE functions the query kernels use (csrc/leafmath.cuh), so the reference's own unit tests for the
leaf arithmetic can be replayed against the GPU path (tests/test_leaf_math.py).  Host numpy arrays in, numpy out.
"""
import numpy as np

from . import _lib
from ._lib import MEM_HOST, check

PRIMAL_TINY = 1e-50


def closest_point(points, triangles, EPS=PRIMAL_TINY, device=0):
    """-> (closest points (n,3), loc (n,) int32)"""
    p = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    t = np.ascontiguousarray(triangles, np.float64).reshape(-1, 9)
    if len(p) != len(t):
        raise ValueError("one triangle per point")
    cp = np.empty_like(p)
    loc = np.empty(len(p), np.int32)
    check(_lib.lib().axb_closest_point_tri(device, p.ctypes.data, t.ctypes.data, len(p), MEM_HOST, float(EPS), cp.ctypes.data, loc.ctypes.data))
    return cp, loc


def squared_distance_point_box(points, boxes, device=0):
    p = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    b = np.ascontiguousarray(boxes, np.float64).reshape(-1, 6)
    if len(p) != len(b):
        raise ValueError("one box per point")
    out = np.empty(len(p), np.float64)
    check(_lib.lib().axb_squared_distance_point_box(device, p.ctypes.data, b.ctypes.data, len(p), MEM_HOST, out.ctypes.data))
    return out


def intersect_ray_box(rays, boxes, tol, normalized=False, device=0):
    """rays (n,6) = origin, direction; normalized=False applies the primal::Ray constructor's normalisation"""
    r = np.ascontiguousarray(rays, np.float64).reshape(-1, 6)
    b = np.ascontiguousarray(boxes, np.float64).reshape(-1, 6)
    if len(r) != len(b):
        raise ValueError("one box per ray")
    out = np.empty(len(r), np.uint8)
    check(_lib.lib().axb_intersect_ray_box(device, r.ctypes.data, b.ctypes.data, len(r), MEM_HOST, int(bool(normalized)), float(tol),
                                           out.ctypes.data))
    return out.astype(bool)


def box_scale(boxes, scale_factor, device=0):
    b = np.ascontiguousarray(boxes, np.float64).reshape(-1, 6)
    out = np.empty_like(b)
    check(_lib.lib().axb_box_scale(device, b.ctypes.data, len(b), MEM_HOST, float(scale_factor), out.ctypes.data))
    return out
