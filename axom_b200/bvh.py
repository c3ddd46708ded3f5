"""Host-side mirror of axom::spin::BVH (spin/BVH.hpp:129-419) over the C ABI.

Method names and argument meaning follow the reference class so the parity tests read like
spin/tests/spin_bvh.cpp.  Inputs may be
  * numpy float64 arrays (host)  -> staged by the library, results come back as numpy arrays;
  * torch CUDA float64 tensors   -> used in place, results come back as torch tensors
    (torch is only the device-memory container here);
  * a tuple/list of per-component 1-D arrays/tensors (the primal::ZipIndexable SoA form).
Everything is executed by libaxb200.so; nothing is computed in Python.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, ArrayDesc, Traverser, check

BVH_BUILD_OK = 0
DEFAULT_SCALE_FACTOR = 1.000123


def _is_torch(a):
    return type(a).__module__.startswith("torch")


class _Keep:
    """descriptor + the arrays it points into (kept alive for the duration of the call)"""

    def __init__(self, desc, refs, count, device):
        self.desc, self.refs, self.count, self.device = desc, refs, count, device


def make_desc(data, ncomp, dtype=np.float64):
    """Build an axb_array_desc for AoS (n, ncomp) or SoA (ncomp x (n,)) input on host or device.
    dtype is the handle's FloatType (float64, or float32 for a BVH created with dtype=np.float32)."""
    d = ArrayDesc()
    d.ncomp = ncomp
    dtype = np.dtype(dtype)
    isz = dtype.itemsize
    tname = "torch." + dtype.name
    if isinstance(data, (tuple, list)):
        if len(data) != ncomp:
            raise ValueError("expected %d component arrays, got %d" % (ncomp, len(data)))
        if all(_is_torch(a) for a in data):
            comps = [a.contiguous() for a in data]
            for a in comps:
                if not a.is_cuda or str(a.dtype) != tname or a.dim() != 1:
                    raise ValueError("SoA components must be 1-D CUDA %s tensors" % dtype.name)
            n = comps[0].numel()
            for c, a in enumerate(comps):
                d.comp[c] = a.data_ptr()
            d.stride_bytes, d.memspace = isz, MEM_DEVICE
            return _Keep(d, comps, n, True)
        comps = [np.ascontiguousarray(a, dtype).reshape(-1) for a in data]
        n = comps[0].size
        for c, a in enumerate(comps):
            if a.size != n:
                raise ValueError("SoA components differ in length")
            d.comp[c] = a.ctypes.data
        d.stride_bytes, d.memspace = isz, MEM_HOST
        return _Keep(d, comps, n, False)
    if _is_torch(data):
        if not data.is_cuda or str(data.dtype) != tname:
            raise ValueError("device input must be a CUDA %s tensor" % dtype.name)
        a = data.contiguous().reshape(-1, ncomp)
        base = a.data_ptr()
        for c in range(ncomp):
            d.comp[c] = base + isz * c
        d.stride_bytes, d.memspace = isz * ncomp, MEM_DEVICE
        return _Keep(d, [a], a.shape[0], True)
    a = np.ascontiguousarray(data, dtype).reshape(-1, ncomp)
    base = a.ctypes.data
    for c in range(ncomp):
        d.comp[c] = base + isz * c
    d.stride_bytes, d.memspace = isz * ncomp, MEM_HOST
    return _Keep(d, [a], a.shape[0], False)


class BVH:
    """spin::BVH<NDIMS, B200_EXEC, FloatType>; dtype = np.float64 (default) or np.float32."""

    def __init__(self, ndims=3, device=0, _borrowed=None, dtype=np.float64):
        self._L = _lib.lib()
        self.ndims = ndims
        self.device = device
        self.dtype = np.dtype(dtype)
        self._owned = _borrowed is None
        if _borrowed is None:
            h = C.c_void_p()
            check(self._L.axb_bvh_create(C.byref(h), ndims, self.dtype.itemsize, device))
            self._h = h
        else:
            self._h = _borrowed

    def __del__(self):
        try:
            if getattr(self, "_owned", False) and getattr(self, "_h", None):
                self._L.axb_bvh_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- configuration (spin/BVH.hpp:255-289) ----
    def setScaleFactor(self, s):
        check(self._L.axb_bvh_set_scale_factor(self._h, float(s)))

    def getScaleFactor(self):
        v = C.c_double()
        check(self._L.axb_bvh_get_scale_factor(self._h, C.byref(v)))
        return v.value

    def setTolerance(self, t):
        check(self._L.axb_bvh_set_tolerance(self._h, float(t)))

    def getTolerance(self):
        v = C.c_double()
        check(self._L.axb_bvh_get_tolerance(self._h, C.byref(v)))
        return v.value

    def setStream(self, cuda_stream_ptr):
        """run on the caller's stream (e.g. torch's current stream): the caller then owns the ordering"""
        check(self._L.axb_bvh_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))
        self._own_stream = False

    def _wait_for_torch(self, k):
        """device inputs produced by torch may still be in flight on torch's stream; the handle works on its own
        stream unless setStream() was called, so wait for them first"""
        if k.device and getattr(self, "_own_stream", True):
            import torch
            torch.cuda.current_stream(self.device).synchronize()

    def setAsync(self, enabled):
        check(self._L.axb_bvh_set_async(self._h, int(bool(enabled))))

    def synchronize(self):
        check(self._L.axb_bvh_synchronize(self._h))

    def writeVtkFile(self, fileName):
        """writeVtkFile(fileName) (spin/BVH.hpp:405): the reference's ASCII VTK dump of the tree, byte for byte"""
        check(self._L.axb_bvh_write_vtk_file(self._h, os.fsencode(fileName)))

    def setProfiling(self, enabled):
        check(self._L.axb_bvh_set_profiling(self._h, int(bool(enabled))))

    def setFindStrategy(self, strategy):
        """0 = one traversal + scatter (default), 1 = count/fill double traversal, 2 = force the overflow path"""
        check(self._L.axb_bvh_set_find_strategy(self._h, int(strategy)))

    def phase_ms(self, name):
        v = C.c_double()
        check(self._L.axb_bvh_get_phase_ms(self._h, name.encode(), C.byref(v)))
        return v.value

    def phases_ms(self, prefix, names):
        """{name: ms} for the phases of the last profiled call that were recorded (a path not taken has none)"""
        out = {}
        for k in names:
            v = C.c_double()
            if self._L.axb_bvh_get_phase_ms(self._h, (prefix + k).encode(), C.byref(v)) == 0:
                out[k] = round(v.value, 4)
        return out

    def launch_count(self):
        v = C.c_int64()
        check(self._L.axb_bvh_launch_count(self._h, C.byref(v)))
        return v.value

    # ---- build (spin/BVH.hpp:424-477) ----
    def initialize(self, boxes, numItems=None):
        k = make_desc(boxes, 2 * self.ndims, self.dtype)
        n = k.count if numItems is None else int(numItems)
        if n > k.count:
            raise ValueError("numItems exceeds the supplied boxes")
        self._wait_for_torch(k)
        st = self._L.axb_bvh_initialize(self._h, C.byref(k.desc), n)
        if st < 0:
            check(st)
        return st

    def isInitialized(self):
        return bool(self._L.axb_bvh_is_initialized(self._h))

    def getBounds(self):
        lo = np.empty(self.ndims, np.float64)
        hi = np.empty(self.ndims, np.float64)
        check(self._L.axb_bvh_get_bounds(self._h, lo.ctypes.data, hi.ctypes.data))
        return lo, hi

    def numLeaves(self):
        v = C.c_int32()
        check(self._L.axb_bvh_num_leaves(self._h, C.byref(v)))
        return v.value

    # ---- queries (spin/BVH.hpp:341-398) ----
    def _find(self, which, k, nq, extra=()):
        self._wait_for_torch(k)
        import_torch = k.device
        total = C.c_int64()
        cand = C.c_void_p()
        if import_torch:
            import torch
            dev = torch.device("cuda", self.device)
            offsets = torch.empty(nq, dtype=torch.int32, device=dev)
            counts = torch.empty(nq, dtype=torch.int32, device=dev)
            po, pc, space = offsets.data_ptr(), counts.data_ptr(), MEM_DEVICE
        else:
            offsets = np.empty(nq, np.int32)
            counts = np.empty(nq, np.int32)
            po, pc, space = offsets.ctypes.data, counts.ctypes.data, MEM_HOST
        fn = getattr(self._L, "axb_bvh_find_" + which)
        check(fn(self._h, C.byref(k.desc), *extra, nq, po, pc, space, C.byref(cand), C.byref(total)))
        t = int(total.value)
        if import_torch:
            import torch
            candidates = torch.empty(t, dtype=torch.int32, device=dev)
            if t:
                # device-to-device copy into a tensor torch owns, then release the library buffer
                _cudart_memcpy_d2d(candidates.data_ptr(), cand.value, 4 * t, self)
        elif t == 0 or not cand.value:
            candidates = np.empty(0, np.int32)  # no query, or no hit: the library hands back a null buffer
        else:
            candidates = np.ctypeslib.as_array(C.cast(cand, C.POINTER(C.c_int32)), shape=(t,)).copy()
        check(self._L.axb_bvh_free_candidates(self._h, cand, space))
        return offsets, counts, candidates

    @staticmethod
    def _count(k, n, what):
        if n is None:
            return k.count
        n = int(n)
        if n < 0 or n > k.count:
            raise ValueError("%s = %d but the array holds %d" % (what, n, k.count))
        return n

    def findPoints(self, points, numPts=None):
        k = make_desc(points, self.ndims, self.dtype)
        return self._find("points", k, self._count(k, numPts, "numPts"))

    def findBoundingBoxes(self, boxes, numBoxes=None):
        k = make_desc(boxes, 2 * self.ndims, self.dtype)
        return self._find("boxes", k, self._count(k, numBoxes, "numBoxes"))

    def findRays(self, origins, directions=None, numRays=None, normalized=False):
        """rays as (origins, directions) AoS pairs or a 2*D-tuple of SoA components.
        normalized=False applies the primal::Ray constructor's normalisation."""
        D = self.ndims
        if directions is None:
            k = make_desc(origins, 2 * D, self.dtype)
        elif _is_torch(origins):
            import torch
            k = make_desc(torch.cat([origins.reshape(-1, D), directions.reshape(-1, D)], dim=1), 2 * D, self.dtype)
        else:
            k = make_desc(np.concatenate([np.asarray(origins, self.dtype).reshape(-1, D),
                                          np.asarray(directions, self.dtype).reshape(-1, D)], axis=1), 2 * D, self.dtype)
        return self._find("rays", k, self._count(k, numRays, "numRays"), extra=(int(bool(normalized)),))

    # ---- traverser / parity views ----
    def getTraverser(self):
        t = Traverser()
        check(self._L.axb_bvh_get_traverser(self._h, C.byref(t)))
        return t

    def arrays(self):
        """host copies of the build artefacts in the reference's layout (parity tests)"""
        n = self.numLeaves()
        inner, D = n - 1, self.ndims
        out = dict(mcodes=np.empty(n, np.uint32), leafs=np.empty(n, np.int32),
                   inner_nodes=np.empty((2 * inner, 2 * D), self.dtype), inner_children=np.empty(2 * inner, np.int32))
        check(self._L.axb_bvh_copy_arrays(self._h, out["mcodes"].ctypes.data, out["leafs"].ctypes.data,
                                          out["inner_nodes"].ctypes.data, out["inner_children"].ctypes.data))
        lo, hi = self.getBounds()
        out["bounds"] = np.concatenate([lo, hi]).astype(self.dtype)
        return out


def _cudart_memcpy_d2d(dst, src, nbytes, bvh):
    import torch
    bvh.synchronize()
    # view the library buffer through the CUDA array interface and let torch copy it
    class _Raw:
        pass
    raw = _Raw()
    raw.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (int(src), False), "version": 3}
    src_t = torch.as_tensor(raw, device=torch.device("cuda", bvh.device))
    dst_t_raw = _Raw()
    dst_t_raw.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (int(dst), False), "version": 3}
    torch.as_tensor(dst_t_raw, device=torch.device("cuda", bvh.device)).copy_(src_t)
    torch.cuda.synchronize(bvh.device)
