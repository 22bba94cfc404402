"""ctypes binding of libtif_b200.so (the C ABI of include/tif_b200.h).

There is no CPU fallback: if the library is missing or fails to load, every product entry point
raises.  Build it in-tree with `python __graft_entry__.py` or `python <pkg>/_build.py`.
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TIF_B200_LIB") or os.path.join(PKG_DIR, "libtif_b200.so")      # the override: A/B variants of _build.build_variant

TIF_OK = 0
ABI_VERSION = 5
BLEND_STRAIGHTFORWARD = 0
BLEND_NORMWARP = 1
BLEND_CUBEMAP_INSIDE = 2
BLEND_CUBEMAP_WARP_ERROR = 3
PLAN_ICOSAHEDRON = 0
PLAN_CUBEMAP = 1
WARP_WRAP = 0
WARP_CONSTANT255 = 1
DTYPE_F32 = 0
DTYPE_F64 = 1
MAX_FACES = 128
MAX_TANGENT_DIM = 2048
IPC_HANDLE_BYTES = 64
PACK_RGB = 0
PACK_GRAY_CV = 1


class TifFace(C.Structure):
    _fields_ = [
        ("theta0", C.c_double), ("phi0", C.c_double),
        ("sin_phi0", C.c_double), ("cos_phi0", C.c_double),
        ("tri", (C.c_double * 2) * 4), ("inner", (C.c_double * 2) * 4),
        ("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
        ("g2p_x", C.c_double), ("g2p_y", C.c_double),
        ("p2g_x", C.c_double), ("p2g_y", C.c_double),
        ("eps_stitch", C.c_double), ("eps_image", C.c_double), ("eps_inner", C.c_double),
        ("col_start", C.c_int32), ("col_stop", C.c_int32),
        ("row_start", C.c_int32), ("row_stop", C.c_int32),
        ("tangent_h", C.c_int32), ("pix_offset", C.c_int32),
        ("n_vertices", C.c_int32), ("n_inner", C.c_int32),
    ]


class TifPlanTables(C.Structure):
    _fields_ = [
        ("gnom_x", C.POINTER(C.c_double)), ("gnom_y", C.POINTER(C.c_double)),
        ("row_sincos", C.POINTER(C.c_double)), ("col_sincos", C.POINTER(C.c_double)),
    ]


class TifPlanInfo(C.Structure):
    _fields_ = [
        ("n_faces", C.c_int32), ("erp_h", C.c_int32), ("erp_w", C.c_int32), ("tangent_w", C.c_int32),
        ("tangent_pixels", C.c_int64), ("max_faces_per_pixel", C.c_int32),
        ("accepted_unique", C.c_int64), ("referenced_tangent_pixels", C.c_int64), ("device_bytes", C.c_int64),
        ("stage1_patches", C.c_int64), ("stage1_staged_patches", C.c_int64),
        ("stage1_tma_patches", C.c_int64), ("stage1_tma_box_rows", C.c_int32), ("stage1_tma_box_bytes", C.c_int32 * 2),
        ("reserved", C.c_int32),
    ]


# every symbol include/tif_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "tif_abi_version": (C.c_int, []),
    "tif_struct_sizes": (C.c_int, [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "tif_status_string": (C.c_char_p, [C.c_int]),
    "tif_last_error": (C.c_char_p, []),
    "tif_plan_create": (C.c_int, [C.POINTER(TifFace), C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(TifPlanTables), C.c_int, _VP, C.POINTER(_VP)]),
    "tif_plan_destroy": (C.c_int, [_VP]),
    "tif_plan_info": (C.c_int, [_VP, C.POINTER(TifPlanInfo)]),
    "tif_plan_export_stage1": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "tif_plan_export_stage2": (C.c_int, [_VP, _VP, _VP, _VP]),
    "tif_erp2ico_u8": (C.c_int, [_VP, _VP, C.c_int, C.c_int, _VP, _VP]),
    "tif_erp2ico_float": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    "tif_pack_tangent_u8": (C.c_int, [_VP, C.c_int64, C.c_int, _VP, _VP]),
    "tif_ico2erp_scratch_bytes": (C.c_int64, [_VP, C.c_int, C.c_int]),
    "tif_ico2erp_flow_f32": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP]),
    "tif_ico2erp_flow_f32_ex": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, _VP, _VP, C.c_int, _VP]),
    "tif_warp_backward_u8": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    "tif_warp_backward_f32": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    "tif_warp_backward_f64": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    "tif_flow_warp_meshgrid": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    "tif_rotation_motion_vector_f64": (C.c_int, [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, _VP, _VP]),
    "tif_flow_cross_covariance_scratch_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "tif_flow_cross_covariance": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP]),
    "tif_blend_weight_points_scratch_bytes": (C.c_int64, [C.c_int64]),
    "tif_blend_weight_points": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP, C.c_int64, C.c_int,
                                          _VP, _VP, _VP]),
    "tif_flow_rotation2d_stats_scratch_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "tif_flow_rotation2d_stats": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP]),
    "tif_flow_rotate_endpoint": (C.c_int, [_VP, C.c_int, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    "tif_flow_error_mats": (C.c_int, [_VP, _VP, C.c_int, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP]),
    "tif_reduce_f64_scratch_bytes": (C.c_int64, [C.c_int]),
    "tif_reduce_f64": (C.c_int, [_VP, C.c_int, C.c_int64, C.c_int, _VP, _VP, _VP]),
    "tif_flow_rad_max": (C.c_int, [_VP, C.c_int, C.c_int64, C.c_int, C.c_double, C.c_double, _VP, _VP]),
    "tif_flow_to_color_u8": (C.c_int, [_VP, C.c_int, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double, _VP, C.c_int, _VP, _VP]),
    "tif_ipc_alloc": (C.c_int, [C.POINTER(_VP), C.c_int64]),
    "tif_ipc_free": (C.c_int, [_VP]),
    "tif_ipc_export": (C.c_int, [_VP, C.c_char_p]),
    "tif_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(_VP)]),
    "tif_ipc_close": (C.c_int, [_VP]),
    "tif_ipc_copy_async": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "tif_erp_of_wraparound": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _VP]),
}

_lib = None


class TifError(RuntimeError):
    pass


def load():
    """Load libtif_b200.so and bind every symbol; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TifError("libtif_b200.so is not built (%s). Run `python __graft_entry__.py` at the repo root; "
                       "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.tif_abi_version() != ABI_VERSION:
        raise TifError("libtif_b200.so ABI version %d, expected %d" % (lib.tif_abi_version(), ABI_VERSION))
    sizes = [C.c_int32(), C.c_int32(), C.c_int32()]
    lib.tif_struct_sizes(*[C.byref(s) for s in sizes])
    expect = (C.sizeof(TifFace), C.sizeof(TifPlanTables), C.sizeof(TifPlanInfo))
    if tuple(s.value for s in sizes) != expect:
        raise TifError("struct layout mismatch between _lib.py %s and libtif_b200.so %s" %
                       (expect, tuple(s.value for s in sizes)))
    _lib = lib
    return lib


def check(status, what):
    if status != TIF_OK:
        lib = load()
        raise TifError("%s failed: %s: %s" % (what, lib.tif_status_string(status).decode(),
                                              lib.tif_last_error().decode()))
