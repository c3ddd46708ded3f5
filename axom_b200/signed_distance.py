"""Host-side mirror of axom::quest::SignedDistance<3> (quest/SignedDistance.hpp:147-397) over the C ABI.

The mint::Mesh argument of the reference is reduced to what SD_GetUcdMeshData extracts
(quest/SignedDistance.cpp:15-45): SoA node coordinates x, y, z and an int32 cells-to-nodes array.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, check
from .bvh import BVH, _is_torch, make_desc


class SignedDistance:
    def __init__(self, x, y, z, cells_to_nodes, nodes_per_cell=3, isWatertight=True, computeSign=True, device=0,
                 cell_node_offsets=None):
        """cell_node_offsets (num_cells+1 int32 offsets into cells_to_nodes) selects a mixed triangle/quad
        mesh (mint::UnstructuredMesh<MIXED_SHAPE>); nodes_per_cell is then ignored."""
        self._L = _lib.lib()
        self.device = device
        self._h = None
        po = None
        if _is_torch(x):
            import torch
            dev = torch.device("cuda", device)
            for name, a, want in (("x", x, torch.float64), ("y", y, torch.float64), ("z", z, torch.float64),
                                  ("cells_to_nodes", cells_to_nodes, torch.int32), ("cell_node_offsets", cell_node_offsets, torch.int32)):
                if a is None:
                    continue
                if not _is_torch(a) or not a.is_cuda or a.device != dev or a.dtype != want:
                    raise TypeError("%s must be a %s tensor on %s (got %s on %s): the library reads the device arrays in place"
                                    % (name, want, dev, getattr(a, "dtype", type(a)), getattr(a, "device", "host")))
            xs = [a.contiguous() for a in (x, y, z)]
            conn = cells_to_nodes.contiguous().reshape(-1)
            torch.cuda.current_stream(xs[0].device).synchronize()  # the library works on its own stream
            px, py, pz, pc = (a.data_ptr() for a in (*xs, conn))
            nn, nc, space = xs[0].numel(), conn.numel() // nodes_per_cell, MEM_DEVICE
            if cell_node_offsets is not None:
                offs = cell_node_offsets.contiguous().reshape(-1)
                po, nc = offs.data_ptr(), offs.numel() - 1
        else:
            xs = [np.ascontiguousarray(a, np.float64).reshape(-1) for a in (x, y, z)]
            conn = np.ascontiguousarray(cells_to_nodes, np.int32).reshape(-1)
            px, py, pz, pc = (a.ctypes.data for a in (*xs, conn))
            nn, nc, space = xs[0].size, conn.size // nodes_per_cell, MEM_HOST
            if cell_node_offsets is not None:
                offs = np.ascontiguousarray(cell_node_offsets, np.int32).reshape(-1)
                po, nc = offs.ctypes.data, offs.size - 1
        h = C.c_void_p()
        check(self._L.axb_sd_create(C.byref(h), device, px, py, pz, nn, pc, po, nc, nodes_per_cell, space,
                                    int(bool(isWatertight)), int(bool(computeSign))))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.axb_sd_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def getBVHTree(self):
        b = C.c_void_p()
        check(self._L.axb_sd_get_bvh(self._h, C.byref(b)))
        return BVH(3, self.device, _borrowed=b)

    def getMeshBounds(self):
        lo, hi = np.empty(3), np.empty(3)
        check(self._L.axb_sd_get_mesh_bounds(self._h, lo.ctypes.data, hi.ctypes.data))
        return lo, hi

    def setMode(self, mode):
        check(self._L.axb_sd_set_mode(self._h, int(mode)))

    def setStream(self, ptr):
        check(self._L.axb_sd_set_stream(self._h, C.c_void_p(ptr)))
        self._own_stream = False

    def setAsync(self, e):
        check(self._L.axb_sd_set_async(self._h, int(bool(e))))

    def synchronize(self):
        check(self._L.axb_sd_synchronize(self._h))

    def setProfiling(self, e):
        check(self._L.axb_sd_set_profiling(self._h, int(e)))

    def phase_ms(self, name):
        v = C.c_double()
        check(self._L.axb_sd_get_phase_ms(self._h, name.encode(), C.byref(v)))
        return v.value

    def launch_count(self):
        v = C.c_int64()
        check(self._L.axb_sd_launch_count(self._h, C.byref(v)))
        return v.value

    def work_counters(self):
        a, b = C.c_int64(), C.c_int64()
        check(self._L.axb_sd_get_work_counters(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def computeDistances(self, queryPts, outClosestPts=False, outNormals=False, out=None):
        """computeDistances(npts, queryPts, outSgnDist, outClosestPts, outNormals) (:527-605).
        Returns (phi, closest_pts|None, normals|None) as numpy arrays (host input) or torch
        tensors (device input).  `out` may supply a preallocated phi tensor/array."""
        k = make_desc(queryPts, 3)
        n = k.count
        if k.device:
            import torch
            if getattr(self, "_own_stream", True):
                torch.cuda.current_stream(self.device).synchronize()  # inputs torch may still be producing
            dev = torch.device("cuda", self.device)
            phi = out if out is not None else torch.empty(n, dtype=torch.float64, device=dev)
            cp = torch.empty((n, 3), dtype=torch.float64, device=dev) if outClosestPts else None
            nr = torch.empty((n, 3), dtype=torch.float64, device=dev) if outNormals else None
            ptr = lambda t: t.data_ptr() if t is not None else None
            space = MEM_DEVICE
        else:
            phi = out if out is not None else np.empty(n, np.float64)
            cp = np.empty((n, 3), np.float64) if outClosestPts else None
            nr = np.empty((n, 3), np.float64) if outNormals else None
            ptr = lambda t: t.ctypes.data if t is not None else None
            space = MEM_HOST
        check(self._L.axb_sd_compute_distances(self._h, C.byref(k.desc), n, ptr(phi), ptr(cp), ptr(nr), space))
        return phi, cp, nr

    def computeDistancesMinReduce(self, comm, queryPts, out=None):
        """BASELINE config C5: this handle holds ONE PART of the surface (computeSign=False), every rank passes the same
        query points, every rank gets the distance to the whole surface (kernel + ncclAllReduce(MIN) on one stream, inside
        the library: axb_sd_compute_distances_minreduce).  comm: axom_b200.comm.Comm."""
        k = make_desc(queryPts, 3)
        n = k.count
        if k.device:
            import torch
            if getattr(self, "_own_stream", True):
                torch.cuda.current_stream(self.device).synchronize()
            d = out if out is not None else torch.empty(n, dtype=torch.float64, device=torch.device("cuda", self.device))
            ptr, space = d.data_ptr(), MEM_DEVICE
        else:
            d = out if out is not None else np.empty(n, np.float64)
            ptr, space = d.ctypes.data, MEM_HOST
        check(self._L.axb_sd_compute_distances_minreduce(self._h, comm._h, C.byref(k.desc), n, ptr, space))
        return d

    def updateMinDistances(self, queryPts, dist):
        """one GPU, partitioned surface: dist (float64 tensor / array, in place) becomes min(dist, distance to this handle's
        part); the entry values bound the search (axb_sd_update_min_distances).  computeSign=False handles."""
        k = make_desc(queryPts, 3)
        n = k.count
        if k.device:
            import torch
            if getattr(self, "_own_stream", True):
                torch.cuda.current_stream(self.device).synchronize()
            assert dist.is_cuda and dist.dtype == torch.float64 and dist.is_contiguous() and dist.numel() == n
            ptr, space = dist.data_ptr(), MEM_DEVICE
        else:
            assert dist.dtype == np.float64 and dist.flags["C_CONTIGUOUS"] and dist.size == n
            ptr, space = dist.ctypes.data, MEM_HOST
        check(self._L.axb_sd_update_min_distances(self._h, C.byref(k.desc), n, ptr, space))
        return dist

    def computeDistance(self, x, y=None, z=0.0):
        """computeDistance(x,y,z) / computeDistance(Point) (:243-266)"""
        p = np.array([[x, y, z]], np.float64) if y is not None else np.asarray(x, np.float64).reshape(1, 3)
        return float(self.computeDistances(p)[0][0])
