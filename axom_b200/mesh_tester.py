"""Host-side mirror of quest::findTriMeshIntersectionsBVH (quest/MeshTester.hpp:67-104) and of
primal::intersect(Triangle3, Triangle3, includeBoundary, EPS) (primal/operators/intersect.hpp:64-71) over the C ABI.

The mint::UnstructuredMesh<SINGLE_SHAPE> argument of the reference is reduced to SoA node coordinates and an
int32 cells-to-nodes array (3 nodes per cell), as for SignedDistance.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, check
from .bvh import BVH, _cudart_memcpy_d2d, _is_torch


class MeshTester:
    """detail::CandidateFinder<AccelType::BVH> (quest/detail/MeshTester_detail.hpp:125-340): construct = initialize()
    (per-cell triangles, AABBs, degenerate flags, BVH build); findTriMeshIntersections() = one fused broad + narrow walk."""

    def __init__(self, x, y, z, cells_to_nodes, device=0):
        self._L = _lib.lib()
        self.device = device
        self._h = None
        if _is_torch(x):
            import torch
            xs = [a.contiguous() for a in (x, y, z)]
            conn = cells_to_nodes.contiguous().reshape(-1)
            torch.cuda.current_stream(xs[0].device).synchronize()  # the library works on its own stream
            px, py, pz, pc = (a.data_ptr() for a in (*xs, conn))
            nn, nc, space = xs[0].numel(), conn.numel() // 3, MEM_DEVICE
        else:
            xs = [np.ascontiguousarray(a, np.float64).reshape(-1) for a in (x, y, z)]
            conn = np.ascontiguousarray(cells_to_nodes, np.int32).reshape(-1)
            px, py, pz, pc = (a.ctypes.data for a in (*xs, conn))
            nn, nc, space = xs[0].size, conn.size // 3, MEM_HOST
        self._torch = space == MEM_DEVICE
        h = C.c_void_p()
        check(self._L.axb_meshtester_create(C.byref(h), device, px, py, pz, nn, pc, nc, space))
        self._h = h
        self.ncells = nc

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.axb_meshtester_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def getBVH(self):
        b = C.c_void_p()
        check(self._L.axb_meshtester_get_bvh(self._h, C.byref(b)))
        return BVH(3, self.device, _borrowed=b)

    def _take(self, p, n, device_out):
        """copy a library-allocated int32 array into numpy / torch and release it"""
        if device_out:
            import torch
            out = torch.empty(n, dtype=torch.int32, device="cuda:%d" % self.device)
            if n:
                _cudart_memcpy_d2d(out.data_ptr(), p.value, 4 * n, self.getBVH())
            check(self._L.axb_meshtester_free(self._h, p, MEM_DEVICE))
            return out
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(max(n, 1),))[:n].copy() if p.value else np.empty(0, np.int32)
        check(self._L.axb_meshtester_free(self._h, p, MEM_HOST))
        return a

    def findTriMeshIntersections(self, intersectionThreshold=1e-8, device_out=None):
        """-> (pairs (n, 2) int32, first < second, in the reference's SEQ order)"""
        dev = self._torch if device_out is None else bool(device_out)
        f, s = C.c_void_p(), C.c_void_p()
        n = C.c_int64()
        check(self._L.axb_meshtester_find_intersections(self._h, float(intersectionThreshold), MEM_DEVICE if dev else MEM_HOST,
                                                        C.byref(f), C.byref(s), C.byref(n)))
        a, b = self._take(f, n.value, dev), self._take(s, n.value, dev)
        if dev:
            import torch
            return torch.stack([a, b], dim=1)
        return np.stack([a, b], axis=1)

    def degenerateIndices(self, device_out=None):
        dev = self._torch if device_out is None else bool(device_out)
        p = C.c_void_p()
        n = C.c_int64()
        check(self._L.axb_meshtester_get_degenerate(self._h, MEM_DEVICE if dev else MEM_HOST, C.byref(p), C.byref(n)))
        return self._take(p, n.value, dev)


def findTriMeshIntersectionsBVH(x, y, z, cells_to_nodes, intersectionThreshold=1e-8, device=0):
    """quest::findTriMeshIntersectionsBVH<ExecSpace, double>(mesh, intersections, degenerateIndices, threshold)
    -> (intersections (n, 2), degenerateIndices)"""
    mt = MeshTester(x, y, z, cells_to_nodes, device=device)
    return mt.findTriMeshIntersections(intersectionThreshold), mt.degenerateIndices()


def intersect_triangles(tris1, tris2, includeBoundary=False, EPS=1e-8, device=0):
    """primal::intersect(t1, t2, includeBoundary, EPS) for n pairs; tris are (n, 3, 3) float64 (numpy or cuda torch)"""
    L = _lib.lib()
    if _is_torch(tris1):
        import torch
        a, b = tris1.contiguous().reshape(-1, 9), tris2.contiguous().reshape(-1, 9)
        out = torch.zeros(a.shape[0], dtype=torch.uint8, device=a.device)
        check(L.axb_tri_tri_intersect(device, a.data_ptr(), b.data_ptr(), a.shape[0], MEM_DEVICE, int(bool(includeBoundary)), float(EPS),
                                      out.data_ptr()))
        return out.bool()
    a = np.ascontiguousarray(tris1, np.float64).reshape(-1, 9)
    b = np.ascontiguousarray(tris2, np.float64).reshape(-1, 9)
    out = np.zeros(a.shape[0], np.uint8)
    check(L.axb_tri_tri_intersect(device, a.ctypes.data, b.ctypes.data, a.shape[0], MEM_HOST, int(bool(includeBoundary)), float(EPS),
                                  out.ctypes.data))
    return out.astype(bool)
