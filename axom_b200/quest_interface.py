"""Host-side mirror of the reference's process-global signed-distance interface
(quest/interface/signed_distance.hpp:88-319; Python hosts of the reference bind the same functions through
quest/interface/python or the Shroud C symbols QUEST_signed_distance_*) plus the input side of the path:
quest::STLReader (quest/readers/STLReader.cpp) and quest::weldTriMeshVertices (quest/MeshTester.cpp:218-333).

Every function forwards to the C symbols of include/axb200_quest.h.  SLIC_ERROR conditions raise QuestError here
(the C default is print + abort, like slic with abort-on-error on).
"""
import ctypes as C
import enum

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST
from .bvh import _is_torch


class SignedDistExec(enum.IntEnum):  # signed_distance.hpp:88-93
    CPU = 0
    OpenMP = 1
    GPU = 2


class QuestError(RuntimeError):
    pass


_HANDLER_T = C.CFUNCTYPE(None, C.c_char_p)
_pending = []


@_HANDLER_T
def _record_error(msg):
    _pending.append(msg.decode())


def _L():
    L = _lib.lib()
    L.axb_quest_set_error_handler(C.cast(_record_error, C.c_void_p))
    return L


def _raise_pending():
    if _pending:
        msg = "; ".join(_pending)
        del _pending[:]
        raise QuestError(msg)


def signed_distance_init(mesh_or_file):
    """signed_distance_init(file) / signed_distance_init(mesh): mesh = (x, y, z, triangles_to_nodes).  Returns 0 / -1."""
    L = _L()
    if isinstance(mesh_or_file, (str, bytes)):
        f = mesh_or_file.encode() if isinstance(mesh_or_file, str) else mesh_or_file
        rc = L.QUEST_signed_distance_init_serial(f)
    else:
        x, y, z, conn = mesh_or_file
        if _is_torch(x):
            xs = [a.contiguous() for a in (x, y, z)]
            c = conn.contiguous().reshape(-1)
            rc = L.axb_quest_signed_distance_init_mesh(*(a.data_ptr() for a in xs), xs[0].numel(), c.data_ptr(), c.numel() // 3, MEM_DEVICE)
        else:
            xs = [np.ascontiguousarray(a, np.float64).reshape(-1) for a in (x, y, z)]
            c = np.ascontiguousarray(conn, np.int32).reshape(-1)
            rc = L.axb_quest_signed_distance_init_mesh(*(a.ctypes.data for a in xs), xs[0].size, c.ctypes.data, c.size // 3, MEM_HOST)
    _raise_pending()
    return rc


def signed_distance_initialized():
    return bool(_L().QUEST_signed_distance_initialized())


def signed_distance_get_mesh_bounds():
    lo, hi = np.empty(3), np.empty(3)
    _L().QUEST_signed_distance_get_mesh_bounds(lo.ctypes.data, hi.ctypes.data)
    _raise_pending()
    return lo, hi


def _setter(name):
    def f(value):
        getattr(_L(), "QUEST_signed_distance_" + name)(value)
        _raise_pending()
    f.__name__ = "signed_distance_" + name
    return f


signed_distance_set_dimension = _setter("set_dimension")
signed_distance_set_closed_surface = _setter("set_closed_surface")
signed_distance_set_compute_signs = _setter("set_compute_signs")
signed_distance_set_allocator = _setter("set_allocator")
signed_distance_set_verbose = _setter("set_verbose")
signed_distance_use_shared_memory = _setter("use_shared_memory")


def signed_distance_set_execution_space(exec_space):
    _L().QUEST_signed_distance_set_execution_space(int(exec_space))
    _raise_pending()


def signed_distance_evaluate(x, y, z, phi=None, with_closest_point=False):
    """scalar x, y, z -> phi (or (phi, cp, normal) with with_closest_point); arrays x, y, z -> phi array
    (signed_distance_evaluate overloads, signed_distance.hpp:231-284)"""
    L = _L()
    if np.isscalar(x):
        if with_closest_point:
            v = [C.c_double() for _ in range(6)]
            p = L.QUEST_signed_distance_evaluate_1(float(x), float(y), float(z), *(C.byref(a) for a in v))
            _raise_pending()
            return p, np.array([a.value for a in v[:3]]), np.array([a.value for a in v[3:]])
        p = L.QUEST_signed_distance_evaluate_0(float(x), float(y), float(z))
        _raise_pending()
        return p
    if _is_torch(x):
        import torch
        xs = [a.contiguous() for a in (x, y, z)]
        out = torch.empty_like(xs[0]) if phi is None else phi
        L.axb_quest_signed_distance_evaluate_n(*(a.data_ptr() for a in xs), xs[0].numel(), out.data_ptr())
        _raise_pending()
        return out
    xs = [np.ascontiguousarray(a, np.float64).reshape(-1) for a in (x, y, z)]
    out = np.empty_like(xs[0]) if phi is None else phi
    L.axb_quest_signed_distance_evaluate_n(*(a.ctypes.data for a in xs), xs[0].size, out.ctypes.data)
    _raise_pending()
    return out


def signed_distance_finalize():
    _L().QUEST_signed_distance_finalize()


def read_stl(file):
    """quest::STLReader: -> (x, y, z, triangles_to_nodes (n, 3) int32); node 3i+k is vertex k of triangle i"""
    L = _L()
    px, py, pz, pc = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    nn, nc = C.c_int32(), C.c_int32()
    if L.axb_stl_read(file.encode(), C.byref(px), C.byref(py), C.byref(pz), C.byref(nn), C.byref(pc), C.byref(nc)) != 0:
        raise QuestError("reading mesh from [%s] failed!" % file)

    def take(p, n, ct, dt):
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(max(n, 1),))[:n].astype(dt, copy=True)
        L.axb_host_free(p)
        return a
    x, y, z = (take(p, nn.value, C.c_double, np.float64) for p in (px, py, pz))
    return x, y, z, take(pc, 3 * nc.value, C.c_int32, np.int32).reshape(-1, 3)


def weldTriMeshVertices(x, y, z, triangles_to_nodes, eps):
    """quest::weldTriMeshVertices(&mesh, eps) -> the welded (x, y, z, triangles_to_nodes)"""
    L = _L()
    xs = [np.array(a, np.float64).reshape(-1) for a in (x, y, z)]
    c = np.array(triangles_to_nodes, np.int32).reshape(-1)
    nn, nc = C.c_int32(xs[0].size), C.c_int32(c.size // 3)
    if L.axb_weld_tri_mesh_vertices(*(a.ctypes.data for a in xs), C.byref(nn), c.ctypes.data, C.byref(nc), float(eps)) != 0:
        raise QuestError("weldTriMeshVertices: bad arguments (eps must be > 0)")
    return xs[0][:nn.value].copy(), xs[1][:nn.value].copy(), xs[2][:nn.value].copy(), c[:3 * nc.value].reshape(-1, 3).copy()
