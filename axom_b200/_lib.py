"""ctypes binding of libaxb200.so (the C ABI in include/axb200.h).

The library is the product: if it is missing this module raises -- there is no Python or
CPU fallback for any compute entry point.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AXB200_LIB") or os.path.join(HERE, "lib", "libaxb200.so")  # AXB200_LIB: tuning builds

AXB_OK = 0
AXB_ERR_BAD_ARG, AXB_ERR_CUDA, AXB_ERR_OVERFLOW, AXB_ERR_NOT_BUILT, AXB_ERR_NO_DEVICE, AXB_ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
MEM_HOST, MEM_DEVICE = 0, 1


class ArrayDesc(C.Structure):
    _fields_ = [("comp", C.c_void_p * 6), ("stride_bytes", C.c_int64), ("ncomp", C.c_int32), ("memspace", C.c_int32)]


class Traverser(C.Structure):
    _fields_ = [("inner_nodes", C.c_void_p), ("inner_node_children", C.c_void_p), ("leaf_nodes", C.c_void_p),
                ("num_leaves", C.c_int32), ("ndims", C.c_int32), ("fp_bytes", C.c_int32)]


class AxbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("%s: %s" % (status_string(status), msg))
        self.status = status


_lib = None

# every symbol include/axb200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_DESC = C.POINTER(ArrayDesc)
_PD = C.POINTER(C.c_double)
SYMBOLS = [
    ("axb_version", C.c_char_p, []),
    ("axb_trim_pool", C.c_int, [C.c_int, C.c_uint64]),
    ("axb_last_error", C.c_char_p, []),
    ("axb_device_count", C.c_int, []),
    ("axb_status_string", C.c_char_p, [C.c_int]),
    ("axb_bvh_create", C.c_int, [_PP, C.c_int, C.c_int, C.c_int]),
    ("axb_bvh_destroy", C.c_int, [_P]),
    ("axb_bvh_set_stream", C.c_int, [_P, _P]),
    ("axb_bvh_set_async", C.c_int, [_P, C.c_int]),
    ("axb_bvh_synchronize", C.c_int, [_P]),
    ("axb_bvh_set_scale_factor", C.c_int, [_P, C.c_double]),
    ("axb_bvh_get_scale_factor", C.c_int, [_P, _PD]),
    ("axb_bvh_set_tolerance", C.c_int, [_P, C.c_double]),
    ("axb_bvh_get_tolerance", C.c_int, [_P, _PD]),
    ("axb_bvh_initialize", C.c_int, [_P, _DESC, C.c_int32]),
    ("axb_bvh_is_initialized", C.c_int, [_P]),
    ("axb_bvh_get_bounds", C.c_int, [_P, _P, _P]),
    ("axb_bvh_get_traverser", C.c_int, [_P, C.POINTER(Traverser)]),
    ("axb_bvh_find_points", C.c_int, [_P, _DESC, C.c_int32, _P, _P, C.c_int, _PP, C.POINTER(C.c_int64)]),
    ("axb_bvh_find_boxes", C.c_int, [_P, _DESC, C.c_int32, _P, _P, C.c_int, _PP, C.POINTER(C.c_int64)]),
    ("axb_bvh_find_rays", C.c_int, [_P, _DESC, C.c_int, C.c_int32, _P, _P, C.c_int, _PP, C.POINTER(C.c_int64)]),
    ("axb_bvh_free_candidates", C.c_int, [_P, _P, C.c_int]),
    ("axb_bvh_set_find_strategy", C.c_int, [_P, C.c_int]),
    ("axb_bvh_num_leaves", C.c_int, [_P, C.POINTER(C.c_int32)]),
    ("axb_bvh_copy_arrays", C.c_int, [_P, _P, _P, _P, _P]),
    ("axb_bvh_write_vtk_file", C.c_int, [_P, C.c_char_p]),
    ("axb_bvh_set_profiling", C.c_int, [_P, C.c_int]),
    ("axb_bvh_get_phase_ms", C.c_int, [_P, C.c_char_p, _PD]),
    ("axb_bvh_launch_count", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("axb_sd_create", C.c_int, [_PP, C.c_int, _P, _P, _P, C.c_int32, _P, _P, C.c_int32, C.c_int32, C.c_int, C.c_int, C.c_int]),
    ("axb_sd_destroy", C.c_int, [_P]),
    ("axb_sd_set_stream", C.c_int, [_P, _P]),
    ("axb_sd_set_async", C.c_int, [_P, C.c_int]),
    ("axb_sd_synchronize", C.c_int, [_P]),
    ("axb_sd_compute_distances", C.c_int, [_P, _DESC, C.c_int32, _P, _P, _P, C.c_int]),
    ("axb_sd_get_bvh", C.c_int, [_P, _PP]),
    ("axb_sd_get_mesh_bounds", C.c_int, [_P, _P, _P]),
    ("axb_sd_set_mode", C.c_int, [_P, C.c_int]),
    ("axb_sd_set_profiling", C.c_int, [_P, C.c_int]),
    ("axb_sd_get_phase_ms", C.c_int, [_P, C.c_char_p, _PD]),
    ("axb_sd_launch_count", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("axb_sd_get_work_counters", C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("axb_meshtester_create", C.c_int, [_PP, C.c_int, _P, _P, _P, C.c_int32, _P, C.c_int32, C.c_int]),
    ("axb_meshtester_destroy", C.c_int, [_P]),
    ("axb_meshtester_find_intersections", C.c_int, [_P, C.c_double, C.c_int, _PP, _PP, C.POINTER(C.c_int64)]),
    ("axb_meshtester_get_degenerate", C.c_int, [_P, C.c_int, _PP, C.POINTER(C.c_int64)]),
    ("axb_meshtester_free", C.c_int, [_P, _P, C.c_int]),
    ("axb_meshtester_get_bvh", C.c_int, [_P, _PP]),
    ("axb_tri_tri_intersect", C.c_int, [C.c_int, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_double, _P]),
    ("axb_closest_point_tri", C.c_int, [C.c_int, _P, _P, C.c_int64, C.c_int, C.c_double, _P, _P]),
    ("axb_squared_distance_point_box", C.c_int, [C.c_int, _P, _P, C.c_int64, C.c_int, _P]),
    ("axb_intersect_ray_box", C.c_int, [C.c_int, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_double, _P]),
    ("axb_box_scale", C.c_int, [C.c_int, _P, C.c_int64, C.c_int, C.c_double, _P]),
    ("axb_dcp_create", C.c_int, [_PP, C.c_int, C.c_int]),
    ("axb_dcp_destroy", C.c_int, [_P]),
    ("axb_dcp_set_object_points", C.c_int, [_P, _P, _P, C.c_int32, C.c_int]),
    ("axb_dcp_generate_bvh_tree", C.c_int, [_P]),
    ("axb_dcp_set_squared_distance_threshold", C.c_int, [_P, C.c_double]),
    ("axb_dcp_set_mode", C.c_int, [_P, C.c_int]),
    ("axb_dcp_get_bvh", C.c_int, [_P, _PP]),
    ("axb_dcp_compute_local_closest_points", C.c_int, [_P, C.c_int, _P, C.c_int32, C.c_int, _P, _P, _P, _P, _P, C.c_int]),
    ("axb_dcp_compute_bounded_closest_points", C.c_int, [_P, C.c_int, _P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    # the exchange steps (NCCL inside the library)
    ("axb_comm_get_unique_id", C.c_int, [_P]),
    ("axb_comm_create", C.c_int, [_PP, C.c_int, C.c_int, _P, C.c_int]),
    ("axb_comm_destroy", C.c_int, [_P]),
    ("axb_comm_get_rank", C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("axb_comm_get_traffic", C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("axb_comm_library", C.c_char_p, []),
    ("axb_comm_allreduce_f64", C.c_int, [_P, _P, C.c_int64, C.c_int, _P]),
    ("axb_sd_compute_distances_minreduce", C.c_int, [_P, _P, _DESC, C.c_int32, _P, C.c_int]),
    ("axb_sd_update_min_distances", C.c_int, [_P, _DESC, C.c_int32, _P, C.c_int]),
    ("axb_dcp_compute_closest_points", C.c_int, [_P, _P, _P, C.c_int32, C.c_int, _P, _P, _P, _P, _P]),
    # quest::MarchingCubes
    ("axb_mc_create", C.c_int, [_PP, C.c_int, C.c_int]),
    ("axb_mc_destroy", C.c_int, [_P]),
    ("axb_mc_set_stream", C.c_int, [_P, _P]),
    ("axb_mc_set_mesh", C.c_int, [_P, _P, C.c_int32, C.c_int]),
    ("axb_mc_set_mask_value", C.c_int, [_P, C.c_int]),
    ("axb_mc_compute_isocontour", C.c_int, [_P, C.c_double]),
    ("axb_mc_get_contour_cell_count", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("axb_mc_get_contour_node_count", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("axb_mc_get_contour_views", C.c_int, [_P, _PP, _PP, _PP, _PP]),
    ("axb_mc_copy_contour", C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    ("axb_mc_clear_output", C.c_int, [_P]),
    ("axb_mc_set_profiling", C.c_int, [_P, C.c_int]),
    ("axb_mc_get_phase_ms", C.c_int, [_P, C.c_char_p, _PD]),
    ("axb_mc_launch_count", C.c_int, [_P, C.POINTER(C.c_int64)]),
    # include/axb200_quest.h: the reference's legacy process-global C surface (wrapQUEST.h:83-127) + STL / welding
    ("QUEST_signed_distance_init_serial", C.c_int, [C.c_char_p]),
    ("QUEST_signed_distance_init_serial_bufferify", C.c_int, [C.c_char_p, C.c_int]),
    ("QUEST_signed_distance_initialized", C.c_bool, []),
    ("QUEST_signed_distance_get_mesh_bounds", None, [_P, _P]),
    ("QUEST_signed_distance_set_dimension", None, [C.c_int]),
    ("QUEST_signed_distance_set_closed_surface", None, [C.c_bool]),
    ("QUEST_signed_distance_set_compute_signs", None, [C.c_bool]),
    ("QUEST_signed_distance_set_allocator", None, [C.c_int]),
    ("QUEST_signed_distance_set_verbose", None, [C.c_bool]),
    ("QUEST_signed_distance_use_shared_memory", None, [C.c_bool]),
    ("QUEST_signed_distance_set_execution_space", None, [C.c_int]),
    ("QUEST_signed_distance_evaluate_0", C.c_double, [C.c_double, C.c_double, C.c_double]),
    ("QUEST_signed_distance_evaluate_1", C.c_double, [C.c_double, C.c_double, C.c_double, _PD, _PD, _PD, _PD, _PD, _PD]),
    ("QUEST_signed_distance_finalize", None, []),
    ("axb_quest_signed_distance_init_mesh", C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, C.c_int]),
    ("axb_quest_signed_distance_evaluate_n", None, [_P, _P, _P, C.c_int, _P]),
    ("axb_stl_read", C.c_int, [C.c_char_p, _PP, _PP, _PP, C.POINTER(C.c_int32), _PP, C.POINTER(C.c_int32)]),
    ("axb_weld_tri_mesh_vertices", C.c_int, [_P, _P, _P, C.POINTER(C.c_int32), _P, C.POINTER(C.c_int32), C.c_double]),
    ("axb_host_free", None, [_P]),
    ("axb_quest_set_error_handler", None, [_P]),
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -m axom_b200.build` (nvcc, sm_100a). "
                              "axom_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def status_string(s):
    return lib().axb_status_string(int(s)).decode()


def check(status):
    if status != AXB_OK:
        raise AxbError(status, lib().axb_last_error().decode())
    return status
