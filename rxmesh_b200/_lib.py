"""ctypes binding of librxmesh_b200.so (the C ABI in include/rxmesh_b200.h).

The product path has no fallback: if the CUDA library is missing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RXM_LIB: another build of the same library (kernel tuning experiments: make NVFLAGS+=-DRXM_BT2=160 OUT=...)
LIB_PATH = os.environ.get("RXM_LIB") or os.path.join(_HERE, "librxmesh_b200.so")

u32p = C.POINTER(C.c_uint32)
u16p = C.POINTER(C.c_uint16)


class PatchView(C.Structure):
    _fields_ = [("patch_id", C.c_uint32), ("n", C.c_uint32 * 3), ("n_owned", C.c_uint32 * 3),
                ("slot_base", C.c_uint32 * 3), ("lin_base", C.c_uint32 * 3),
                ("packed", C.c_uint32), ("ev", u16p), ("fe", u16p), ("fv", u16p),
                ("voff_e", u16p), ("voff_f", u16p), ("eoff_f", u16p), ("fan_off", u16p), ("fan_v", u16p),
                ("fan_f", u16p), ("fan_total", C.c_uint32), ("owner", u32p * 3), ("stash", u32p),
                ("n_stash", C.c_uint32), ("ltog", u32p * 3), ("ff", u16p), ("ef", u16p), ("fan_e", u16p),
                ("r2_idx", u16p), ("r2_off", u16p), ("r2_val", u16p), ("ext_owner", u32p),
                ("n_r2", C.c_uint32), ("n_ext", C.c_uint32), ("r2_total", C.c_uint32)]


class PatcherFile(C.Structure):
    _fields_ = [("header", C.c_uint32 * 9), ("vec", u32p * 7), ("len", C.c_uint64 * 7),
                ("patching_time_ms", C.c_float)]


# every symbol declared in include/rxmesh_b200.h: (restype, argtypes)
SYMBOLS = {
    "rxm_last_error": (C.c_char_p, []),
    "rxm_version": (C.c_char_p, []),
    "rxm_init": (C.c_int, [C.c_int]),
    "rxm_mesh_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int,
                                  C.POINTER(C.c_void_p)]),
    "rxm_mesh_create_ex": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32,
                                     C.POINTER(C.c_void_p)]),
    "rxm_mesh_to_device": (C.c_int, [C.c_void_p]),
    "rxm_mesh_compact": (C.c_int, [C.c_void_p]),
    "rxm_mesh_destroy": (None, [C.c_void_p]),
    "rxm_mesh_info": (C.c_uint64, [C.c_void_p, C.c_int]),
    "rxm_mesh_build_seconds": (C.c_double, [C.c_void_p, C.c_int]),
    "rxm_mesh_patch": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(PatchView)]),
    "rxm_mesh_slot_to_global": (u32p, [C.c_void_p, C.c_int]),
    "rxm_mesh_global_to_slot": (u32p, [C.c_void_p, C.c_int]),
    "rxm_mesh_elem_patch": (u32p, [C.c_void_p, C.c_int]),
    "rxm_mesh_slot_base": (u32p, [C.c_void_p, C.c_int]),
    "rxm_mesh_lin_base": (u32p, [C.c_void_p, C.c_int]),
    "rxm_mesh_edges": (u32p, [C.c_void_p]),
    "rxm_mesh_face_edges": (u32p, [C.c_void_p]),
    "rxm_mesh_device_slot_base": (C.c_void_p, [C.c_void_p, C.c_int]),
    "rxm_mesh_device_lin_base": (C.c_void_p, [C.c_void_p, C.c_int]),
    "rxm_mesh_view": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "rxm_mesh_launch_box": (C.c_int, [C.c_void_p, C.c_int, u32p, u32p, u32p]),
    "rxm_attr_create": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                  C.POINTER(C.c_void_p)]),
    "rxm_attr_destroy": (None, [C.c_void_p]),
    "rxm_attr_release": (C.c_int, [C.c_void_p, C.c_int]),
    "rxm_attr_data": (C.c_void_p, [C.c_void_p, C.c_int]),
    "rxm_attr_count": (C.c_uint64, [C.c_void_p]),
    "rxm_attr_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "rxm_attr_move": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rxm_attr_copy_from": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rxm_attr_upload_global": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_attr_download_global": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_attr_from_global_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_attr_to_global_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_query_store": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_query_consume": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_vertex_normals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "rxm_laplacian_smooth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint32,
                                       C.c_void_p]),
    "rxm_bilateral_filter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "rxm_bilateral_deferred": (C.c_uint64, [C.c_void_p]),
    "rxm_mcf_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_uint32, C.c_float, C.c_float,
                                C.c_void_p, C.c_void_p]),
    "rxm_mcf_solve_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_uint32, C.c_float,
                                   C.c_float, C.c_void_p, C.c_void_p]),
    "rxm_query_csr": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                C.POINTER(C.c_uint64), C.c_void_p]),
    "rxm_boundary_vertices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_vertex_normals_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_laplacian_smooth_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                            C.c_uint32, C.c_void_p]),
    "rxm_query_consume_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_mesh_set_active_patches": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "rxm_mesh_halo_slots": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(u32p),
                                      C.POINTER(C.c_uint64)]),
    "rxm_free": (None, [C.c_void_p]),
    "rxm_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "rxm_attr_gather_slots": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "rxm_attr_scatter_slots": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "rxm_attr_push_slots": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "rxm_mesh_pipe_plan": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_fused_halo_create": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "rxm_fused_halo_flags": (C.c_void_p, [C.c_void_p]),
    "rxm_fused_halo_sync_blocks": (C.c_uint32, [C.c_void_p]),
    "rxm_fused_halo_set": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "rxm_fused_halo_destroy": (None, [C.c_void_p]),
    "rxm_laplacian_smooth_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int,
                                             C.c_uint32, C.c_void_p]),
    "rxm_multi_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p)]),
    "rxm_multi_destroy": (None, [C.c_void_p]),
    "rxm_multi_info": (C.c_uint64, [C.c_void_p, C.c_int, C.c_int]),
    "rxm_multi_shard_mesh": (C.c_void_p, [C.c_void_p, C.c_int]),
    "rxm_multi_laplacian_smooth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint32]),
    "rxm_multi_vertex_normals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rxm_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rxm_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "rxm_ipc_close": (C.c_int, [C.c_void_p]),
    "rxm_attr_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_double),
                                  C.POINTER(C.c_uint64), C.c_void_p]),
    "rxm_patcher_file_read": (C.c_int, [C.c_char_p, C.POINTER(PatcherFile)]),
    "rxm_patcher_file_free": (None, [C.POINTER(PatcherFile)]),
    "rxm_mesh_save_patcher_file": (C.c_int, [C.c_void_p, C.c_char_p]),
    "rxm_set_async": (None, [C.c_int]),
    "rxm_stream_sync": (C.c_int, [C.c_void_p]),
    "rxm_launch_count": (C.c_uint64, []),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (make -C rxmesh_b200/csrc). rxmesh_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


class RXMeshError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise RXMeshError(lib().rxm_last_error().decode())
