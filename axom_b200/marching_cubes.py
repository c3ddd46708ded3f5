"""Host-side mirror of quest::MarchingCubes (quest/MarchingCubes.hpp:107-306) over the C ABI (axb_mc_*).

The reference reads a multi-domain Conduit Blueprint mesh.  Conduit is a tree of named nodes; here the same tree is a
plain dict of dicts whose leaves are 1-D numpy arrays (host) or CUDA torch tensors (device, used in place) -- the paths
read are exactly the ones MeshViewUtil reads (quest/MeshViewUtil.hpp:454-482,560-607,803-848):

    mesh[<domain>]["topologies"][<topo>]["coordset"]                      name of the coordset
    mesh[<domain>]["topologies"][<topo>]["elements"]["dims"]["i","j","k"] real cells per direction
                                                   ...["dims"]["offsets"] ghost offsets of the coordinates  (optional)
                                                   ...["dims"]["strides"] strides of the coordinates        (optional)
    mesh[<domain>]["coordsets"][<cs>]["values"]["x","y","z"]              one array per direction (not interleaved)
    mesh[<domain>]["fields"][<f>]["values" | "strides" | "offsets" | "association"]
    mesh[<domain>]["state"]["domain_id"]                                                                    (optional)

domain_views() reduces every domain to the ghost-free strided views MarchingCubesImpl holds; those go through the C ABI.
"""
import ctypes as C
import enum

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, check


class MarchingCubesDataParallelism(enum.IntEnum):
    """quest/MarchingCubes.hpp:49-54.  Both variants give the same contour; the device path does not distinguish them."""
    byPolicy = 0
    hybridParallel = 1
    fullParallel = 2


class DomainView:
    """the views of one domain: flat base arrays + element offset + element strides (ArrayView::subspan(offsets, realShape))"""
    __slots__ = ("cell_shape", "coords", "coords_offset", "coords_strides", "fcn", "fcn_offset", "fcn_strides", "mask",
                 "mask_offset", "mask_strides", "domain_id")


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _ints(v, ndim, default):
    if v is None:
        return list(default)
    return [int(x) for x in (v.tolist() if hasattr(v, "tolist") else v)][:ndim]


def _field_view(dom, name, cell_shape, ndim):
    """getConstFieldView(name, withGhosts=false) (MeshViewUtil.hpp:560-607) -> (values, element offset, strides, on_vertex)"""
    f = dom["fields"][name]
    assoc = f["association"]
    if assoc not in ("vertex", "element"):
        raise ValueError("MeshViewUtil only supports vertex and element-based fields")
    on_vertex = 1 if assoc == "vertex" else 0
    default, t = [], 1
    for d in range(ndim):
        default.append(t)
        t *= cell_shape[d] + on_vertex
    strides = _ints(f.get("strides"), ndim, default)
    offsets = _ints(f.get("offsets"), ndim, [0] * ndim)
    return f["values"], sum(o * s for o, s in zip(offsets, strides)), strides, on_vertex


def domain_views(bpMesh, topologyName, fcnField, maskField=""):
    """one DomainView per child of the multi-domain mesh, in child order (MarchingCubes.cpp:76-80)"""
    if "coordsets" in bpMesh or "topologies" in bpMesh:  # conduit::blueprint::mesh::is_multi_domain (MarchingCubes.cpp:51-52)
        raise ValueError("MarchingCubes class input mesh must be in multidomain format.")
    views = []
    for pos, (_, dom) in enumerate(bpMesh.items()):
        topo = dom["topologies"][topologyName]
        if topo.get("type", "structured") != "structured":
            raise ValueError("MarchingCubes needs a structured topology")
        dims = topo["elements"]["dims"]
        names = [k for k in ("i", "j", "k") if k in dims]
        ndim = len(names)
        v = DomainView()
        v.cell_shape = [int(dims[k]) for k in names]
        cs = dom["coordsets"][topo["coordset"]]["values"]
        v.coords = [cs[k] for k in ("x", "y", "z")[:ndim]]
        default, t = [], 1
        for d in range(ndim):  # computeCoordsDataLayout (:824-834): direction 0 fastest
            default.append(t)
            t *= v.cell_shape[d] + 1
        v.coords_strides = _ints(dims.get("strides"), ndim, default)
        off = _ints(dims.get("offsets"), ndim, [0] * ndim)
        v.coords_offset = sum(o * s for o, s in zip(off, v.coords_strides))
        v.fcn, v.fcn_offset, v.fcn_strides, on_vertex = _field_view(dom, fcnField, v.cell_shape, ndim)
        if not on_vertex:
            raise ValueError("the function field must be vertex-associated")
        if maskField:
            v.mask, v.mask_offset, v.mask_strides, mv = _field_view(dom, maskField, v.cell_shape, ndim)
            if mv:
                raise ValueError("the mask field must be element-associated")
        else:
            v.mask, v.mask_offset, v.mask_strides = None, 0, [0] * ndim
        st = dom.get("state", {})
        v.domain_id = int(st["domain_id"]) if "domain_id" in st else pos  # getDomainId (MarchingCubesSingleDomain.cpp:168-176)
        views.append(v)
    return views


def slab_domain(field_slab, next_first_plane, axes, k0, domain_id, fcn="phi", topology="mesh"):
    """A Blueprint domain for one z-slab of a nodal field on a rectilinear lattice -- how a field that is SHARDED across
    ranks by contiguous z-slabs is contoured without gathering it: `field_slab` holds the rank's node planes k0, k0+1, ...
    (x fastest, then y, then z) and `next_first_plane` the first node plane of the next rank's slab (None for the last
    rank) -- the halo the cells between the two slabs need, one plane per rank and the only data exchanged.
    axes = (ax, ay, az): the lattice coordinates per direction.  numpy arrays or torch tensors (host or CUDA) alike.
    The slabs' contours, taken in rank order with parent ids offset by k0 * (cells per plane), are the single-domain
    contour of the whole field bit for bit (tests/test_marching_cubes.py, bench.py at N > 1).
    -> ({name: domain}, number of cell planes in the slab)"""
    ax, ay, az = axes
    nx, ny = len(ax), len(ay)
    plane = nx * ny
    if _is_torch(field_slab):
        import torch
        vals = torch.cat([field_slab.reshape(-1), next_first_plane.reshape(-1)]) if next_first_plane is not None else field_slab.reshape(-1)
        nplanes = vals.numel() // plane
        zz, yy, xx = torch.meshgrid(az[k0:k0 + nplanes], ay, ax, indexing="ij")
        coords = {"x": xx.reshape(-1).contiguous(), "y": yy.reshape(-1).contiguous(), "z": zz.reshape(-1).contiguous()}
    else:
        vals = np.concatenate([np.ravel(field_slab), np.ravel(next_first_plane)]) if next_first_plane is not None else np.ravel(field_slab)
        nplanes = vals.size // plane
        zz, yy, xx = np.meshgrid(np.asarray(az)[k0:k0 + nplanes], ay, ax, indexing="ij")
        coords = {"x": np.ascontiguousarray(xx.ravel()), "y": np.ascontiguousarray(yy.ravel()), "z": np.ascontiguousarray(zz.ravel())}
    dom = {"coordsets": {"coords": {"type": "explicit", "values": coords}},
           "topologies": {topology: {"type": "structured", "coordset": "coords",
                                     "elements": {"dims": {"i": nx - 1, "j": ny - 1, "k": nplanes - 1}}}},
           "fields": {fcn: {"association": "vertex", "topology": topology, "values": vals}},
           "state": {"domain_id": int(domain_id)}}
    return {"domain_%06d" % int(domain_id): dom}, nplanes - 1


class McDomain(C.Structure):
    """axb_mc_domain (include/axb200.h)"""
    _fields_ = [("cell_shape", C.c_int64 * 3), ("coords", C.c_void_p * 3), ("coords_strides", C.c_int64 * 3), ("fcn", C.c_void_p),
                ("fcn_strides", C.c_int64 * 3), ("mask", C.c_void_p), ("mask_strides", C.c_int64 * 3), ("domain_id", C.c_int64)]


def pack_domains(views, keep):
    """DomainView list -> (McDomain array, memspace).  `keep` receives every array whose pointer is used."""
    arr = (McDomain * max(len(views), 1))()
    device = None

    def ptr(a, dtype, torch_dtype_name, offset, itemsize):
        nonlocal device
        if _is_torch(a) and not a.is_cuda:
            a = a.numpy()  # a (pinned) host tensor is a host array
        if _is_torch(a):
            import torch
            want = getattr(torch, torch_dtype_name)
            if a.dtype != want:
                raise TypeError("device arrays must be CUDA tensors of dtype %s" % torch_dtype_name)
            a = a.contiguous().reshape(-1)
            is_dev, p = True, a.data_ptr()
        else:
            a = np.ascontiguousarray(a, dtype).reshape(-1)
            is_dev, p = False, a.ctypes.data
        if device is None:
            device = is_dev
        elif device != is_dev:
            raise TypeError("all mesh arrays must live in the same memory space")
        keep.append(a)
        return p + offset * itemsize

    for k, v in enumerate(views):
        d = arr[k]
        nd = len(v.cell_shape)
        for i in range(3):
            d.cell_shape[i] = v.cell_shape[i] if i < nd else 1
            d.coords_strides[i] = v.coords_strides[i] if i < nd else 0
            d.fcn_strides[i] = v.fcn_strides[i] if i < nd else 0
            d.mask_strides[i] = v.mask_strides[i] if i < nd else 0
            d.coords[i] = ptr(v.coords[i], np.float64, "float64", v.coords_offset, 8) if i < nd else None
        d.fcn = ptr(v.fcn, np.float64, "float64", v.fcn_offset, 8)
        d.mask = ptr(v.mask, np.int32, "int32", v.mask_offset, 4) if v.mask is not None else None
        d.domain_id = v.domain_id
    return arr, (MEM_DEVICE if device else MEM_HOST)


class MarchingCubes:
    """quest::MarchingCubes.  runtimePolicy / allocatorID are accepted for signature parity: the only execution space here is
    the B200 (`device` = CUDA ordinal)."""

    def __init__(self, runtimePolicy="cuda", allocatorID=None, dataParallelism=MarchingCubesDataParallelism.byPolicy, device=0):
        self._L = _lib.lib()
        self.device = device
        self._h = None
        self._ndims = None
        self._mesh = None
        self._topology = ""
        self._mask_field = ""
        self._fcn_field = ""
        self._mask_val = 1
        self._dirty = False
        self._keep = []
        self._cache = None  # (device_out, arrays) of the last read-out; dropped by compute / clear
        self.dataParallelism = MarchingCubesDataParallelism(dataParallelism)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.axb_mc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- input ---------------------------------------------------------------------------
    def setMesh(self, bpMesh, topologyName, maskField=""):
        self._mesh, self._topology, self._mask_field = bpMesh, topologyName, maskField
        self._dirty = True

    def setFunctionField(self, fcnField):
        self._fcn_field = fcnField
        self._dirty = True

    def setMaskValue(self, maskVal):
        self._mask_val = int(maskVal)

    def _push_mesh(self):
        if self._mesh is None or not self._fcn_field:
            raise RuntimeError("MarchingCubes: setMesh() and setFunctionField() must precede computeIsocontour()")
        views = domain_views(self._mesh, self._topology, self._fcn_field, self._mask_field)
        ndims = len(views[0].cell_shape) if views else (self._ndims or 3)
        if any(len(v.cell_shape) != ndims for v in views):
            raise ValueError("all domains must have the same dimension")
        if self._h is None or ndims != self._ndims:
            if self._h is not None:
                if self.getContourCellCount():
                    raise ValueError("cannot change the dimension while a contour is held; call clearOutput() first")
                self._L.axb_mc_destroy(self._h)
            h = C.c_void_p()
            check(self._L.axb_mc_create(C.byref(h), ndims, self.device))
            self._h, self._ndims = h, ndims
        keep = []
        arr, space = pack_domains(views, keep)
        if space == MEM_DEVICE:
            import torch
            torch.cuda.current_stream(self.device).synchronize()  # the library works on its own stream
        check(self._L.axb_mc_set_mesh(self._h, arr, len(views), space))
        self._keep = keep if space == MEM_DEVICE else []  # device arrays are used in place: keep them alive
        # The reference reads its views of the field and mask live on every computeIsocontour (m_fcnView / m_maskView).
        # Device arrays are read in place here, so they behave the same; HOST arrays are staged by set_mesh, so the
        # staging is repeated before every contour (an in-place update of the field between contours is then seen).
        self._dirty = space != MEM_DEVICE

    # -- compute -------------------------------------------------------------------------
    def computeIsocontour(self, contourVal=0.0):
        """adds the contour at contourVal to the contour mesh computed so far (MarchingCubes.cpp:107-147)"""
        if self._dirty or self._h is None:
            self._push_mesh()
        elif self._keep:
            # device-resident inputs are read in place on the library's own stream: whatever torch has queued on its
            # current stream (an in-place update of the field) must have finished first
            import torch
            torch.cuda.current_stream(self.device).synchronize()
        self._cache = None
        check(self._L.axb_mc_set_mask_value(self._h, self._mask_val))
        check(self._L.axb_mc_compute_isocontour(self._h, float(contourVal)))

    def clearOutput(self):
        self._cache = None
        if self._h is not None:
            check(self._L.axb_mc_clear_output(self._h))

    # -- output --------------------------------------------------------------------------
    def getContourCellCount(self):
        if self._h is None:
            return 0
        n = C.c_int64()
        check(self._L.axb_mc_get_contour_cell_count(self._h, C.byref(n)))
        return n.value

    getContourFacetCount = getContourCellCount

    def getContourNodeCount(self):
        if self._h is None:
            return 0
        n = C.c_int64()
        check(self._L.axb_mc_get_contour_node_count(self._h, C.byref(n)))
        return n.value

    def _contour(self, device_out):
        if self._cache is not None and self._cache[0] == bool(device_out):
            return self._cache[1]
        out = self._read_contour(device_out)
        self._cache = (bool(device_out), out)
        return out

    def _read_contour(self, device_out):
        n, D = self.getContourCellCount(), self._ndims or 3
        if device_out:
            import torch
            dev = "cuda:%d" % self.device
            ids = torch.empty((n, D), dtype=torch.int32, device=dev)
            xyz = torch.empty((n * D, D), dtype=torch.float64, device=dev)
            par = torch.empty(n, dtype=torch.int32, device=dev)
            dom = torch.empty(n, dtype=torch.int32, device=dev)
            if n:
                check(self._L.axb_mc_copy_contour(self._h, MEM_DEVICE, ids.data_ptr(), xyz.data_ptr(), par.data_ptr(), dom.data_ptr()))
            return ids, xyz, par, dom
        ids = np.empty((n, D), np.int32)
        xyz = np.empty((n * D, D), np.float64)
        par = np.empty(n, np.int32)
        dom = np.empty(n, np.int32)
        if n:
            check(self._L.axb_mc_copy_contour(self._h, MEM_HOST, ids.ctypes.data, xyz.ctypes.data, par.ctypes.data, dom.ctypes.data))
        return ids, xyz, par, dom

    def getContourFacetCorners(self, device_out=False):
        return self._contour(device_out)[0]

    def getContourNodeCoords(self, device_out=False):
        return self._contour(device_out)[1]

    def getContourFacetParents(self, device_out=False):
        return self._contour(device_out)[2]

    def getContourFacetDomainIds(self, device_out=False):
        return self._contour(device_out)[3]

    def relinquishContourData(self, device_out=False):
        """-> (facetNodeIds, facetNodeCoords, facetParentIds, facetDomainIds); the object then holds no contour (:271-285)"""
        out = self._contour(device_out)
        self.clearOutput()
        return out

    def populateContourMesh(self, cellIdField="", domainIdField=""):
        """the mint::UnstructuredMesh<SINGLE_SHAPE> of the reference as a dict of host arrays (:169-233): nodes, cells and the
        two optional cell-centred fields"""
        ids, xyz, par, dom = self._contour(False)
        mesh = {"dimension": self._ndims or 3, "cell_type": "SEGMENT" if (self._ndims or 3) == 2 else "TRIANGLE", "nodes": xyz,
                "cells": ids, "fields": {}}
        if cellIdField:
            mesh["fields"][cellIdField] = par
        if domainIdField:
            mesh["fields"][domainIdField] = dom
        return mesh

    # -- measurement ---------------------------------------------------------------------
    def set_profiling(self, on=True):
        if self._h is None:
            self._push_mesh()
        check(self._L.axb_mc_set_profiling(self._h, int(bool(on))))

    def phase_ms(self, name):
        ms = C.c_double()
        check(self._L.axb_mc_get_phase_ms(self._h, name.encode(), C.byref(ms)))
        return ms.value

    def launch_count(self):
        if self._h is None:
            return 0
        n = C.c_int64()
        check(self._L.axb_mc_launch_count(self._h, C.byref(n)))
        return n.value
