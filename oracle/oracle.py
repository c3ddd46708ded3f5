"""ctypes wrapper over the CPU oracle libraries.  TEST INFRASTRUCTURE ONLY.

Two back ends share one interface (same entry points, different prefix):
  kind="port"       oracle/liboracle.so          -- our restatement (axb_oracle.cpp)
  kind="reference"  oracle/_ref/libaxom_ref.so   -- the real LLNL/axom SEQ_EXEC path (ref_driver.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (axom_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(HERE, "liboracle.so")
_PORT_F32 = os.path.join(HERE, "liboracle_f32.so")
_REF = os.path.join(HERE, "_ref", "libaxom_ref.so")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build_port():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so", "liboracle_f32.so"])


def have_reference():
    return os.path.exists(_REF)


class _Lib:
    """kind: "port" | "reference" (FloatType = double), "port_f32" | "reference_f32" (FloatType = float, BVH only)"""

    def __init__(self, kind):
        self.kind = kind
        self.f32 = kind.endswith("_f32")
        if kind in ("port", "port_f32"):
            path = _PORT_F32 if self.f32 else _PORT
            if not os.path.exists(path):
                build_port()
            self.lib = C.CDLL(path)
            self.pfx = "axof_" if self.f32 else "axo_"
        elif kind in ("reference", "reference_f32"):
            if not os.path.exists(_REF):
                raise FileNotFoundError(_REF + " (run python oracle/build_ref.py where /root/reference exists)")
            self.lib = C.CDLL(_REF)
            self.pfx = "axreff_" if self.f32 else "axref_"
        else:
            raise ValueError(kind)
        f = self.fn
        f("bvh_create").restype = C.c_void_p
        f("bvh_create").argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double]
        f("bvh_destroy").argtypes = [C.c_void_p]
        f("bvh_num_leaves").argtypes = [C.c_void_p]
        f("bvh_get").argtypes = [C.c_void_p] + [C.c_void_p] * 8
        f("free").argtypes = [C.c_void_p]
        for name in ("bvh_find_points", "bvh_find_boxes"):
            f(name).restype = C.c_int64
            f(name).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        f("bvh_find_rays").restype = C.c_int64
        f("bvh_find_rays").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_void_p)]
        if not self.f32:
            f("bvh_count_points_omp").restype = C.c_int64
            f("bvh_count_points_omp").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            f("bvh_count_boxes_omp").restype = C.c_int64
            f("bvh_count_boxes_omp").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            f("bvh_count_rays_omp").restype = C.c_int64
            f("bvh_count_rays_omp").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            f("sd_create").restype = C.c_void_p
            f("sd_create").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
            f("sd_create_mixed").restype = C.c_void_p
            f("sd_create_mixed").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
            f("sd_destroy").argtypes = [C.c_void_p]
            f("sd_compute").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
            f("dcp_create").restype = C.c_void_p
            f("dcp_create").argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
            f("dcp_destroy").argtypes = [C.c_void_p, C.c_int]
            f("dcp_compute_local").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            f("tri_tri_intersect").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
            f("closest_point_tri").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
            f("squared_distance_point_box").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            f("intersect_ray_box").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
            f("box_scale").argtypes = [C.c_void_p, C.c_int, C.c_double]
            f("find_tri_mesh_intersections").restype = C.c_int64
            f("find_tri_mesh_intersections").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                                         C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                         C.POINTER(C.c_int64)]
        f("max_threads").restype = C.c_int

    def fn(self, name):
        return getattr(self.lib, self.pfx + name)


_libs = {}


def lib(kind="port"):
    if kind not in _libs:
        _libs[kind] = _Lib(kind)
    return _libs[kind]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Bvh:
    """spin::BVH<D, SEQ_EXEC, double> on the CPU (port or real reference)."""

    def __init__(self, boxes, ndims=3, scale=-1.0, tol=-1.0, kind="port"):
        self.L = lib(kind)
        self.ndims = ndims
        self.ft = np.float32 if self.L.f32 else np.float64
        boxes = np.ascontiguousarray(boxes, dtype=self.ft).reshape(-1, 2 * ndims)
        self.n_in = boxes.shape[0]
        self.h = self.L.fn("bvh_create")(ndims, _ptr(boxes), self.n_in, float(scale), float(tol))
        self.n = self.L.fn("bvh_num_leaves")(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.fn("bvh_destroy")(self.h)
                self.h = None
        except Exception:
            pass

    def arrays(self):
        n, inner, D = self.n, self.n - 1, self.ndims
        out = dict(
            mcodes=np.empty(n, np.uint32), leafs=np.empty(n, np.int32), lchild=np.empty(inner, np.int32),
            rchild=np.empty(inner, np.int32), parents=np.empty(inner + n, np.int32),
            inner_nodes=np.empty((2 * inner, 2 * D), self.ft), inner_children=np.empty(2 * inner, np.int32),
            bounds=np.empty(2 * D, self.ft))
        self.L.fn("bvh_get")(self.h, *[_ptr(out[k]) for k in
                                       ("mcodes", "leafs", "lchild", "rchild", "parents", "inner_nodes", "inner_children", "bounds")])
        return out

    def write_vtk(self, file_name):
        """BVH::writeVtkFile of the unmodified reference (kind "reference" / "reference_f32" only)"""
        f = self.L.fn("bvh_write_vtk")
        f.argtypes = [C.c_void_p, C.c_char_p]
        f.restype = None
        f(self.h, os.fsencode(file_name))

    def _collect(self, total, cand_p):
        cand = np.ctypeslib.as_array(C.cast(cand_p, C.POINTER(C.c_int32)), shape=(max(int(total), 1),))[:int(total)].copy()
        self.L.fn("free")(cand_p)
        return cand

    def find_points(self, pts):
        pts = np.ascontiguousarray(pts, self.ft).reshape(-1, self.ndims)
        q = pts.shape[0]
        off, cnt = np.empty(q, np.int32), np.empty(q, np.int32)
        cp = C.c_void_p()
        tot = self.L.fn("bvh_find_points")(self.h, _ptr(pts), q, _ptr(off), _ptr(cnt), C.byref(cp))
        return off, cnt, self._collect(tot, cp)

    def find_boxes(self, qboxes):
        qb = np.ascontiguousarray(qboxes, self.ft).reshape(-1, 2 * self.ndims)
        q = qb.shape[0]
        off, cnt = np.empty(q, np.int32), np.empty(q, np.int32)
        cp = C.c_void_p()
        tot = self.L.fn("bvh_find_boxes")(self.h, _ptr(qb), q, _ptr(off), _ptr(cnt), C.byref(cp))
        return off, cnt, self._collect(tot, cp)

    def find_rays(self, origins, dirs, normalize=True):
        o = np.ascontiguousarray(origins, self.ft).reshape(-1, self.ndims)
        d = np.ascontiguousarray(dirs, self.ft).reshape(-1, self.ndims)
        q = o.shape[0]
        off, cnt = np.empty(q, np.int32), np.empty(q, np.int32)
        cp = C.c_void_p()
        tot = self.L.fn("bvh_find_rays")(self.h, _ptr(o), _ptr(d), q, int(bool(normalize)), _ptr(off), _ptr(cnt), C.byref(cp))
        return off, cnt, self._collect(tot, cp)

    def count_points_omp(self, pts, nthreads=0):
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, self.ndims)
        cnt = np.empty(pts.shape[0], np.int32)
        tot = self.L.fn("bvh_count_points_omp")(self.h, _ptr(pts), pts.shape[0], _ptr(cnt), int(nthreads))
        return int(tot), cnt


    def count_boxes_omp(self, boxes, nthreads=0):
        b = np.ascontiguousarray(boxes, np.float64).reshape(-1, 2 * self.ndims)
        cnt = np.empty(b.shape[0], np.int32)
        tot = self.L.fn("bvh_count_boxes_omp")(self.h, _ptr(b), b.shape[0], _ptr(cnt), int(nthreads))
        return int(tot), cnt

    def count_rays_omp(self, origins, directions, nthreads=0):
        """directions are normalised, as the primal::Ray constructor does"""
        o = np.ascontiguousarray(origins, np.float64).reshape(-1, self.ndims)
        d = np.ascontiguousarray(directions, np.float64).reshape(-1, self.ndims)
        cnt = np.empty(o.shape[0], np.int32)
        tot = self.L.fn("bvh_count_rays_omp")(self.h, _ptr(o), _ptr(d), o.shape[0], _ptr(cnt), int(nthreads))
        return int(tot), cnt


class SignedDistance:
    """quest::SignedDistance<3, SEQ_EXEC> on the CPU (port or real reference)."""

    def __init__(self, x, y, z, conn, nodes_per_cell=3, watertight=True, compute_sign=True, kind="port", offsets=None):
        """offsets (ncells+1 int32 into conn) selects a mixed triangle/quad mesh (UnstructuredMesh<MIXED_SHAPE>)"""
        self.L = lib(kind)
        self.x = np.ascontiguousarray(x, np.float64)
        self.y = np.ascontiguousarray(y, np.float64)
        self.z = np.ascontiguousarray(z, np.float64)
        self.conn = np.ascontiguousarray(conn, np.int32).reshape(-1)
        if offsets is not None:
            self.offsets = np.ascontiguousarray(offsets, np.int32)
            self.h = self.L.fn("sd_create_mixed")(_ptr(self.x), _ptr(self.y), _ptr(self.z), self.x.size, _ptr(self.conn),
                                                  _ptr(self.offsets), self.offsets.size - 1, int(watertight), int(compute_sign))
            return
        ncells = self.conn.size // nodes_per_cell
        self.h = self.L.fn("sd_create")(_ptr(self.x), _ptr(self.y), _ptr(self.z), self.x.size, _ptr(self.conn), ncells,
                                        nodes_per_cell, int(watertight), int(compute_sign))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.fn("sd_destroy")(self.h)
                self.h = None
        except Exception:
            pass

    def compute(self, qpts, want_cp=False, want_normals=False, nthreads=1):
        q = np.ascontiguousarray(qpts, np.float64).reshape(-1, 3)
        n = q.shape[0]
        phi = np.empty(n, np.float64)
        cp = np.empty((n, 3), np.float64) if want_cp else None
        nr = np.empty((n, 3), np.float64) if want_normals else None
        self.L.fn("sd_compute")(self.h, _ptr(q), n, _ptr(phi), _ptr(cp), _ptr(nr), int(nthreads))
        return phi, cp, nr


class DistributedClosestPointRank:
    """one rank of quest::DistributedClosestPoint: flattened object points + domain ids, BVH over BoxType{pt}
    (generateBVHTreeImpl), and computeLocalClosestPoints on caller-held state arrays."""

    def __init__(self, points, domain_ids=None, ndims=3, kind="port"):
        self.L = lib(kind)
        self.ndims = ndims
        self.pts = np.ascontiguousarray(points, np.float64).reshape(-1, ndims)
        n = self.pts.shape[0]
        self.dom = np.zeros(n, np.int32) if domain_ids is None else np.ascontiguousarray(domain_ids, np.int32)
        self.h = self.L.fn("dcp_create")(ndims, _ptr(self.pts), _ptr(self.dom), n)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.fn("dcp_destroy")(self.h, self.ndims)
                self.h = None
        except Exception:
            pass

    def compute_local(self, rank, query, state=None, sq_threshold=np.finfo(np.float64).max):
        """state = dict(cp_index, cp_domain_index, cp_rank, cp_coords, cp_distance) updated in place; None = is_first"""
        q = np.ascontiguousarray(query, np.float64).reshape(-1, self.ndims)
        n = q.shape[0]
        first = state is None
        if first:
            state = {"cp_index": np.empty(n, np.int32), "cp_domain_index": np.empty(n, np.int32), "cp_rank": np.empty(n, np.int32),
                     "cp_coords": np.empty((n, self.ndims), np.float64), "cp_distance": np.empty(n, np.float64)}
        self.L.fn("dcp_compute_local")(self.h, self.ndims, int(rank), float(sq_threshold), _ptr(q), n, int(first), _ptr(state["cp_index"]),
                                       _ptr(state["cp_domain_index"]), _ptr(state["cp_rank"]), _ptr(state["cp_coords"]),
                                       _ptr(state["cp_distance"]))
        return state


def tri_tri_intersect(tris1, tris2, include_boundary=False, eps=1e-8, kind="port"):
    """primal::intersect(Triangle3, Triangle3, includeBoundary, EPS) on n pairs; tris are (n, 3, 3) doubles"""
    a = np.ascontiguousarray(tris1, np.float64).reshape(-1, 9)
    b = np.ascontiguousarray(tris2, np.float64).reshape(-1, 9)
    out = np.zeros(a.shape[0], np.uint8)
    lib(kind).fn("tri_tri_intersect")(_ptr(a), _ptr(b), a.shape[0], int(include_boundary), float(eps), _ptr(out))
    return out.astype(bool)


def closest_point_tri(points, triangles, eps=1e-50, kind="port"):
    """primal::closest_point(Point, Triangle, &loc, EPS) on n items -> (cp (n,3), loc (n,))"""
    p = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    t = np.ascontiguousarray(triangles, np.float64).reshape(-1, 9)
    cp, loc = np.empty_like(p), np.empty(len(p), np.int32)
    lib(kind).fn("closest_point_tri")(_ptr(p), _ptr(t), len(p), float(eps), _ptr(cp), _ptr(loc))
    return cp, loc


def squared_distance_point_box(points, boxes, kind="port"):
    p = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    b = np.ascontiguousarray(boxes, np.float64).reshape(-1, 6)
    out = np.empty(len(p), np.float64)
    lib(kind).fn("squared_distance_point_box")(_ptr(p), _ptr(b), len(p), _ptr(out))
    return out


def intersect_ray_box(rays, boxes, tol, kind="port"):
    """primal::detail::intersect_ray(Ray(origin, direction), box, ip, tol); the Ray constructor normalises"""
    r = np.ascontiguousarray(rays, np.float64).reshape(-1, 6)
    b = np.ascontiguousarray(boxes, np.float64).reshape(-1, 6)
    out = np.zeros(len(r), np.uint8)
    lib(kind).fn("intersect_ray_box")(_ptr(r), _ptr(b), len(r), 1, float(tol), _ptr(out))
    return out.astype(bool)


def box_scale(boxes, scale, kind="port"):
    b = np.ascontiguousarray(boxes, np.float64).reshape(-1, 6).copy()
    lib(kind).fn("box_scale")(_ptr(b), len(b), float(scale))
    return b


def find_tri_mesh_intersections(x, y, z, conn, threshold=1e-8, kind="port"):
    """quest::findTriMeshIntersectionsBVH<SEQ_EXEC,double>: ((npairs, 2) int32 pairs in the SEQ order, degenerate ids)"""
    L = lib(kind)
    x, y, z = (np.ascontiguousarray(a, np.float64) for a in (x, y, z))
    conn = np.ascontiguousarray(conn, np.int32).reshape(-1, 3)
    f, s, d = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nd = C.c_int64()
    n = L.fn("find_tri_mesh_intersections")(_ptr(x), _ptr(y), _ptr(z), x.size, _ptr(conn), conn.shape[0], float(threshold),
                                            C.byref(f), C.byref(s), C.byref(d), C.byref(nd))

    def take(p, k):
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(max(k, 1),))[:k].copy()
        L.fn("free")(p)
        return a
    return np.stack([take(f, n), take(s, n)], axis=1), take(d, nd.value)


def ref_legacy_signed_distance(stl_file, x, y, z, closed_surface=True, compute_sign=True):
    """the REAL reference's process-global interface end to end (quest::signed_distance_init(file), batched evaluate,
    get_mesh_bounds, finalize).  Reference only.  -> (phi, lo, hi)"""
    L = lib("reference").lib
    x, y, z = (np.ascontiguousarray(a, np.float64) for a in (x, y, z))
    phi, lo, hi = np.empty_like(x), np.empty(3), np.empty(3)
    L.axref_legacy_signed_distance.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
    rc = L.axref_legacy_signed_distance(stl_file.encode(), int(closed_surface), int(compute_sign), _ptr(x), _ptr(y), _ptr(z), x.size,
                                        _ptr(phi), _ptr(lo), _ptr(hi))
    if rc != 0:
        raise RuntimeError("reference signed_distance_init failed")
    return phi, lo, hi


def ref_stl_read_weld(stl_file, eps=0.0):
    """the REAL reference's quest::STLReader (+ quest::weldTriMeshVertices when eps > 0).  -> (x, y, z, conn (n,3))"""
    L = lib("reference").lib
    px, py, pz, pc = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    nn, nc = C.c_int32(), C.c_int32()
    L.axref_stl_read_weld.argtypes = [C.c_char_p, C.c_double] + [C.POINTER(C.c_void_p)] * 3 + [C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                                                                                             C.POINTER(C.c_int32)]
    if L.axref_stl_read_weld(stl_file.encode(), float(eps), C.byref(px), C.byref(py), C.byref(pz), C.byref(nn), C.byref(pc),
                             C.byref(nc)) != 0:
        raise RuntimeError("reference STL read failed")

    def take(p, n, ct):
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(max(n, 1),))[:n].copy()
        L.axref_free(p)
        return a
    x, y, z = (take(p, nn.value, C.c_double) for p in (px, py, pz))
    return x, y, z, take(pc, 3 * nc.value, C.c_int32).reshape(-1, 3)


def max_threads(kind="port"):
    return lib(kind).fn("max_threads")()


# ---- quest::MarchingCubes (mc_oracle.cpp) --------------------------------------------------------------------------
_MC = os.path.join(HERE, "liboracle_mc.so")
_mc_lib = None


class _McDomain(C.Structure):  # mc_oracle.cpp: McDomain
    _fields_ = [("cell_shape", C.c_int64 * 3), ("coords", C.c_void_p * 3), ("coords_strides", C.c_int64 * 3), ("fcn", C.c_void_p),
                ("fcn_strides", C.c_int64 * 3), ("mask", C.c_void_p), ("mask_strides", C.c_int64 * 3), ("domain_id", C.c_int64)]


def mc_lib():
    global _mc_lib
    if _mc_lib is None:
        if not os.path.exists(_MC):
            subprocess.check_call(["make", "-s", "-C", HERE, "liboracle_mc.so"])
        L = C.CDLL(_MC)
        L.axo_mc_compute_isocontour.restype = C.c_int64
        L.axo_mc_compute_isocontour.argtypes = [C.c_int, C.c_void_p, C.c_int32, C.c_double, C.c_int, C.c_int64] + [C.POINTER(C.c_void_p)] * 4
        L.axo_mc_free.argtypes = [C.c_void_p]
        L.axo_mc_table.argtypes = [C.c_int] * 3
        L.axo_mc_num_contour_cells.argtypes = [C.c_int] * 2
        L.axo_mc_mapping.argtypes = [C.c_int] + [C.c_void_p] * 4
        L.axo_mc_to_multi_index.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _mc_lib = L
    return _mc_lib


def mc_isocontour(views, contour_val=0.0, mask_val=1, first_facet=0):
    """the restated MarchingCubes::computeIsocontour over host views (objects with the attributes of
    axom_b200.marching_cubes.DomainView holding numpy arrays).  -> (facetNodeIds (n, D), nodeCoords (n*D, D), parents, domainIds)
    for the facets this call adds to a contour that already holds `first_facet` facets."""
    L = mc_lib()
    nd = len(views[0].cell_shape) if views else 3
    arr = (_McDomain * max(len(views), 1))()
    keep = []

    def ptr(a, dt, off):
        a = np.ascontiguousarray(a, dt).reshape(-1)
        keep.append(a)
        return a.ctypes.data + off * a.itemsize

    for k, v in enumerate(views):
        d = arr[k]
        for i in range(3):
            d.cell_shape[i] = v.cell_shape[i] if i < nd else 1
            d.coords_strides[i] = v.coords_strides[i] if i < nd else 0
            d.fcn_strides[i] = v.fcn_strides[i] if i < nd else 0
            d.mask_strides[i] = v.mask_strides[i] if i < nd else 0
            d.coords[i] = ptr(v.coords[i], np.float64, v.coords_offset) if i < nd else None
        d.fcn = ptr(v.fcn, np.float64, v.fcn_offset)
        d.mask = ptr(v.mask, np.int32, v.mask_offset) if v.mask is not None else None
        d.domain_id = v.domain_id
    p = [C.c_void_p() for _ in range(4)]
    n = L.axo_mc_compute_isocontour(nd, arr, len(views), float(contour_val), int(mask_val), int(first_facet), *[C.byref(x) for x in p])

    def take(q, count, ct, shape):
        a = np.ctypeslib.as_array(C.cast(q, C.POINTER(ct)), shape=(max(count, 1),))[:count].copy().reshape(shape)
        L.axo_mc_free(q)
        return a
    return (take(p[0], n * nd, C.c_int32, (n, nd)), take(p[1], n * nd * nd, C.c_double, (n * nd, nd)), take(p[2], n, C.c_int32, (n,)),
            take(p[3], n, C.c_int32, (n,)))


def ref_mc_isocontour(bpMesh, topology, fcn_field, mask_field="", mask_val=1, contour_vals=(0.0,), data_parallelism=0):
    """the REAL reference's quest::MarchingCubes (seq policy) on a Blueprint-shaped dict tree (numpy leaves), through the
    Conduit mock of oracle/conduit_stub.  Every contour value is computed in turn into the same (accumulating) contour mesh.
    -> (facetNodeIds (n, D), nodeCoords (n*D, D), parents, domainIds)"""
    L = lib("reference").lib
    L.axref_node_new.restype = C.c_void_p
    L.axref_node_free.argtypes = [C.c_void_p]
    L.axref_node_set_string.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.axref_node_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
    L.axref_node_set_external.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int64]
    L.axref_mc_run.restype = C.c_int64
    L.axref_mc_run.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)] + \
        [C.POINTER(C.c_void_p)] * 4
    root = L.axref_node_new()
    keep = []

    def walk(n, path):
        if isinstance(n, dict):
            for k, v in n.items():
                walk(v, path + "/" + k if path else k)
        elif isinstance(n, str):
            L.axref_node_set_string(root, path.encode(), n.encode())
        elif isinstance(n, (int, np.integer)):
            L.axref_node_set_int(root, path.encode(), int(n))
        else:
            a = np.ascontiguousarray(n)
            kind = {np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.float64): 2}[a.dtype]
            keep.append(a)
            L.axref_node_set_external(root, path.encode(), kind, a.ctypes.data, a.size)
    walk(bpMesh, "")
    cv = np.ascontiguousarray(contour_vals, np.float64)
    p = [C.c_void_p() for _ in range(4)]
    nd = C.c_int()
    try:
        n = L.axref_mc_run(root, topology.encode(), fcn_field.encode(), mask_field.encode(), int(mask_val), cv.ctypes.data, cv.size,
                           int(data_parallelism), C.byref(nd), *[C.byref(x) for x in p])
    finally:
        L.axref_node_free(root)
    D = nd.value

    def take(q, count, ct, shape):
        a = np.ctypeslib.as_array(C.cast(q, C.POINTER(ct)), shape=(max(count, 1),))[:count].copy().reshape(shape)
        L.axref_free(q)
        return a
    return (take(p[0], n * D, C.c_int32, (n, D)), take(p[1], n * D * D, C.c_double, (n * D, D)), take(p[2], n, C.c_int32, (n,)),
            take(p[3], n, C.c_int32, (n,)))


def ref_mc_table(dim, case, edge):
    return lib("reference").lib.axref_mc_table(dim, case, edge)


def ref_mc_num_contour_cells(dim, case):
    return lib("reference").lib.axref_mc_num_contour_cells(dim, case)
