"""PyTorch custom ops (`torch.ops.tangentflow.*`) over the C ABI of libtif_b200.so.

Torch is plumbing only: tensors own the device memory, `torch.cuda.current_stream()` provides the
stream, and the ops make the kernels addressable from torch programs.  Every op takes CUDA tensors
and launches hand-written sm_100a kernels through ctypes; there is no eager / CPU implementation
registered, so calling one with CPU tensors raises.
"""
import itertools
import weakref

import torch

from . import _geometry, _lib

# op handle -> Plan (custom ops cannot take Python objects).  Handles are a monotonically increasing counter, never
# reused, and the registry holds the plans weakly: dropping the last strong reference (`_geometry.clear_plans()`)
# lets Plan.__del__ run tif_plan_destroy and return the plan's HBM.
_registry = weakref.WeakValueDictionary()
_next_handle = itertools.count(1)


def register_plan(plan):
    handle = getattr(plan, "op_handle", None)
    if handle is None:
        handle = plan.op_handle = next(_next_handle)
    _registry[handle] = plan
    return handle


def unregister_all():
    _registry.clear()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.TifError("tangentflow ops need CUDA tensors (no CPU fallback)")


_libdef = torch.library.Library("tangentflow", "DEF")
_libdef.define("erp2ico(Tensor erp, int plan, bool full_face_image) -> Tensor")
_libdef.define("ico2erp_flow(Tensor tangent_flow, Tensor? erp_src, Tensor? erp_tar, int plan, int blend, "
               "bool wrap_around) -> Tensor")
_libdef.define("warp_backward(Tensor image, Tensor flow, int mode) -> Tensor")
_libdef.define("flow_warp_meshgrid(Tensor flow_u, Tensor flow_v) -> Tensor")
_libdef.define("erp_of_wraparound_(Tensor(a!) erp_flow, float u_threshold) -> Tensor(a!)")


def erp2ico_out(p, erp, full_face_image, out):
    """tif_erp2ico_u8 on the current stream into a caller-owned [B,P,4] uint8 tensor."""
    _need_cuda(erp, out)
    if erp.dtype != torch.uint8 or erp.dim() != 4 or erp.shape[1:] != (p.erp_h, p.erp_w, 3) or not erp.is_contiguous():
        raise ValueError("erp2ico expects contiguous uint8 [B,%d,%d,3], got %s %s" % (p.erp_h, p.erp_w, erp.dtype, tuple(erp.shape)))
    if out.dtype != torch.uint8 or tuple(out.shape) != (erp.shape[0], p.tangent_pixels, 4) or not out.is_contiguous():
        raise ValueError("erp2ico output must be contiguous uint8 [B,%d,4]" % p.tangent_pixels)
    with torch.cuda.device(erp.device):
        _lib.check(p.lib.tif_erp2ico_u8(p.handle, erp.data_ptr(), erp.shape[0], int(full_face_image), out.data_ptr(),
                                        _stream()), "tif_erp2ico_u8")
    return out


def _erp2ico(erp, plan, full_face_image):
    p = _registry[plan]
    _need_cuda(erp)
    erp = erp.contiguous()
    out = torch.empty((erp.shape[0], p.tangent_pixels, 4), dtype=torch.uint8, device=erp.device)
    return erp2ico_out(p, erp, full_face_image, out)


def ico2erp_flow_out(p, tangent_flow, erp_src, erp_tar, blend, wrap_around, out, scratch=None, pass_mask=3):
    """tif_ico2erp_flow_f32 on the current stream into a caller-owned [B,H,W,2] float32 tensor, or -- `out` given as an
    int -- to a raw device address with room for it (a slice of a peer-mapped sharding.PeerGatherBuffer).
    Inputs must be contiguous; `scratch` (uint8, >= tif_ico2erp_scratch_bytes) is allocated here when omitted."""
    out_is_ptr = isinstance(out, int)
    _need_cuda(tangent_flow, erp_src, erp_tar, None if out_is_ptr else out)
    batch = tangent_flow.shape[0]
    if tangent_flow.dtype != torch.float32 or tuple(tangent_flow.shape) != (batch, p.tangent_pixels, 2) or not tangent_flow.is_contiguous():
        raise ValueError("ico2erp_flow expects contiguous float32 [B,%d,2]" % p.tangent_pixels)
    if not out_is_ptr and (out.dtype != torch.float32 or tuple(out.shape) != (batch, p.erp_h, p.erp_w, 2) or not out.is_contiguous()):
        raise ValueError("ico2erp_flow output must be contiguous float32 [B,%d,%d,2]" % (p.erp_h, p.erp_w))
    src_ptr = tar_ptr = None
    if blend in (_lib.BLEND_NORMWARP, _lib.BLEND_CUBEMAP_WARP_ERROR):
        for name, img in (("erp_src", erp_src), ("erp_tar", erp_tar)):
            if img is None or img.dtype != torch.uint8 or tuple(img.shape) != (batch, p.erp_h, p.erp_w, 3) or not img.is_contiguous():
                raise ValueError("this blend needs contiguous uint8 %s [B,%d,%d,3]" % (name, p.erp_h, p.erp_w))
        src_ptr, tar_ptr = erp_src.data_ptr(), erp_tar.data_ptr()
    need = int(p.lib.tif_ico2erp_scratch_bytes(p.handle, batch, blend))
    if scratch is None:
        scratch = torch.empty(need, dtype=torch.uint8, device=tangent_flow.device)
    elif scratch.numel() * scratch.element_size() < need:
        raise ValueError("ico2erp_flow needs %d bytes of scratch" % need)
    scratch_ptr = scratch.data_ptr()
    with torch.cuda.device(tangent_flow.device):
        _lib.check(p.lib.tif_ico2erp_flow_f32_ex(p.handle, tangent_flow.data_ptr(), src_ptr, tar_ptr, batch, int(blend),
                                                 int(wrap_around), out if out_is_ptr else out.data_ptr(), scratch_ptr,
                                                 int(pass_mask), _stream()),
                   "tif_ico2erp_flow_f32")
    return out


def _ico2erp_flow(tangent_flow, erp_src, erp_tar, plan, blend, wrap_around):
    p = _registry[plan]
    _need_cuda(tangent_flow, erp_src, erp_tar)
    tangent_flow = tangent_flow.contiguous()
    if blend in (_lib.BLEND_NORMWARP, _lib.BLEND_CUBEMAP_WARP_ERROR):
        if erp_src is None or erp_tar is None:
            raise ValueError("this blend needs erp_src and erp_tar")
        erp_src, erp_tar = erp_src.contiguous(), erp_tar.contiguous()
    out = torch.empty((tangent_flow.shape[0], p.erp_h, p.erp_w, 2), dtype=torch.float32, device=tangent_flow.device)
    return ico2erp_flow_out(p, tangent_flow, erp_src, erp_tar, blend, wrap_around, out)


def _flow_dtype(flow):
    if flow.dtype == torch.float32:
        return _lib.DTYPE_F32
    if flow.dtype == torch.float64:
        return _lib.DTYPE_F64
    raise ValueError("flow must be float32 or float64, got %s" % flow.dtype)


def _warp_backward(image, flow, mode):
    _need_cuda(image, flow)
    if image.dim() != 4 or flow.dim() != 4 or flow.shape[-1] != 2 or flow.shape[:3] != image.shape[:3]:
        raise ValueError("warp_backward expects image [B,H,W,C] and flow [B,H,W,2]")
    image, flow = image.contiguous(), flow.contiguous()
    b, h, w, ch = image.shape
    out = torch.empty_like(image)
    lib = _lib.load()
    with torch.cuda.device(image.device):
        if image.dtype == torch.uint8:
            fn, name = lib.tif_warp_backward_u8, "tif_warp_backward_u8"
        elif image.dtype == torch.float32:
            fn, name = lib.tif_warp_backward_f32, "tif_warp_backward_f32"
        elif image.dtype == torch.float64:
            fn, name = lib.tif_warp_backward_f64, "tif_warp_backward_f64"
        else:
            raise ValueError("warp_backward image must be uint8, float32 or float64, got %s" % image.dtype)
        _lib.check(fn(image.data_ptr(), flow.data_ptr(), _flow_dtype(flow), b, h, w, ch, int(mode), out.data_ptr(),
                      _stream()), name)
    return out


def _flow_warp_meshgrid(flow_u, flow_v):
    _need_cuda(flow_u, flow_v)
    if flow_u.shape != flow_v.shape or flow_u.dim() != 3 or flow_u.dtype != flow_v.dtype:
        raise ValueError("flow_warp_meshgrid expects two [B,H,W] tensors of one dtype")
    flow_u, flow_v = flow_u.contiguous(), flow_v.contiguous()
    b, h, w = flow_u.shape
    out = torch.empty((b, 2, h, w), dtype=flow_u.dtype, device=flow_u.device)
    lib = _lib.load()
    with torch.cuda.device(flow_u.device):
        _lib.check(lib.tif_flow_warp_meshgrid(flow_u.data_ptr(), flow_v.data_ptr(), _flow_dtype(flow_u), b, h, w,
                                              out.data_ptr(), _stream()), "tif_flow_warp_meshgrid")
    return out


def _erp_of_wraparound_(erp_flow, u_threshold):
    _need_cuda(erp_flow)
    if erp_flow.dim() != 4 or erp_flow.shape[-1] != 2 or not erp_flow.is_contiguous():
        raise ValueError("erp_of_wraparound_ expects a contiguous [B,H,W,2] tensor")
    b, h, w, _ = erp_flow.shape
    lib = _lib.load()
    with torch.cuda.device(erp_flow.device):
        _lib.check(lib.tif_erp_of_wraparound(erp_flow.data_ptr(), _flow_dtype(erp_flow), b, h, w, float(u_threshold),
                                             _stream()), "tif_erp_of_wraparound")
    return erp_flow


def rotation_motion_vector(rotation, height, width, wraparound, device):
    """tif_rotation_motion_vector_f64: [H,W,2] float64 CUDA tensor for a 3x3 rotation (host numpy / list)."""
    import ctypes as C
    import numpy as np
    rot = np.ascontiguousarray(np.asarray(rotation, dtype=np.float64).reshape(9))
    out = torch.empty((int(height), int(width), 2), dtype=torch.float64, device=device)
    lib = _lib.load()
    with torch.cuda.device(out.device):
        _lib.check(lib.tif_rotation_motion_vector_f64(rot.ctypes.data_as(C.POINTER(C.c_double)), int(height), int(width),
                                                      int(bool(wraparound)), out.data_ptr(), _stream()),
                   "tif_rotation_motion_vector_f64")
    return out


def flow_cross_covariance(flow):
    """tif_flow_cross_covariance: [B,9] float64 CUDA tensor for a [B,H,W,2] flow."""
    _need_cuda(flow)
    if flow.dim() != 4 or flow.shape[-1] != 2:
        raise ValueError("flow_cross_covariance expects [B,H,W,2]")
    flow = flow.contiguous()
    b, h, w, _ = flow.shape
    lib = _lib.load()
    out = torch.empty((b, 9), dtype=torch.float64, device=flow.device)
    scratch = torch.empty(max(int(lib.tif_flow_cross_covariance_scratch_bytes(b, h, w)), 8), dtype=torch.uint8, device=flow.device)
    with torch.cuda.device(flow.device):
        _lib.check(lib.tif_flow_cross_covariance(flow.data_ptr(), _flow_dtype(flow), b, h, w, out.data_ptr(), scratch.data_ptr(),
                                                 _stream()), "tif_flow_cross_covariance")
    return out


_IMAGE_DTYPES = {torch.float32: 0, torch.float64: 1, torch.uint8: 2}


def blend_weight_points(image_src, image_tar, x_src, y_src, x_tar, y_tar, weight_kind):
    """tif_blend_weight_points: [N] float64 weights for N source / target positions (float64 CUDA tensors) and two
    [H,W,C] images of one dtype (uint8, float32 or float64)."""
    _need_cuda(image_src, image_tar, x_src, y_src, x_tar, y_tar)
    if image_src.dtype != image_tar.dtype or image_src.dtype not in _IMAGE_DTYPES or image_src.shape != image_tar.shape:
        raise ValueError("blend_weight_points: the two images must share shape and dtype (uint8, float32 or float64)")
    image_src, image_tar = image_src.contiguous(), image_tar.contiguous()
    h, w, c = image_src.shape
    coords = [t.to(torch.float64).contiguous().reshape(-1) for t in (x_src, y_src, x_tar, y_tar)]
    n = coords[0].numel()
    if any(t.numel() != n for t in coords):
        raise ValueError("blend_weight_points: coordinate arrays differ in length")
    lib = _lib.load()
    out = torch.empty((n,), dtype=torch.float64, device=image_src.device)
    if n == 0:
        return out
    scratch = torch.empty(max(int(lib.tif_blend_weight_points_scratch_bytes(n)), 8), dtype=torch.uint8, device=image_src.device)
    with torch.cuda.device(image_src.device):
        _lib.check(lib.tif_blend_weight_points(image_src.data_ptr(), image_tar.data_ptr(), _IMAGE_DTYPES[image_src.dtype], h, w, c,
                                               coords[0].data_ptr(), coords[1].data_ptr(), coords[2].data_ptr(),
                                               coords[3].data_ptr(), n, int(weight_kind), out.data_ptr(), scratch.data_ptr(),
                                               _stream()), "tif_blend_weight_points")
    return out


def flow_rotation2d_stats(flow, col_start, col_stop):
    """tif_flow_rotation2d_stats: [B,6] float64 CUDA tensor (sum u; sign / positive / negative sums of v in the band)."""
    _need_cuda(flow)
    if flow.dim() != 4 or flow.shape[-1] != 2:
        raise ValueError("flow_rotation2d_stats expects [B,H,W,2]")
    flow = flow.contiguous()
    b, h, w, _ = flow.shape
    lib = _lib.load()
    out = torch.empty((b, 6), dtype=torch.float64, device=flow.device)
    scratch = torch.empty(max(int(lib.tif_flow_rotation2d_stats_scratch_bytes(b, h, w)), 8), dtype=torch.uint8, device=flow.device)
    with torch.cuda.device(flow.device):
        _lib.check(lib.tif_flow_rotation2d_stats(flow.data_ptr(), _flow_dtype(flow), b, h, w, int(col_start), int(col_stop),
                                                 out.data_ptr(), scratch.data_ptr(), _stream()), "tif_flow_rotation2d_stats")
    return out


def flow_rotate_endpoint(flow, rotation_field, wraparound):
    """tif_flow_rotate_endpoint: flow [B,H,W,2] (float32/64), rotation_field [H,W,2] float64 -> like flow."""
    _need_cuda(flow, rotation_field)
    flow, rotation_field = flow.contiguous(), rotation_field.contiguous()
    b, h, w, _ = flow.shape
    if rotation_field.dtype != torch.float64 or tuple(rotation_field.shape) != (h, w, 2):
        raise ValueError("rotation_field must be float64 [H,W,2]")
    out = torch.empty_like(flow)
    lib = _lib.load()
    with torch.cuda.device(flow.device):
        _lib.check(lib.tif_flow_rotate_endpoint(flow.data_ptr(), _flow_dtype(flow), rotation_field.data_ptr(), b, h, w,
                                                int(bool(wraparound)), out.data_ptr(), _stream()), "tif_flow_rotate_endpoint")
    return out


_libimpl = torch.library.Library("tangentflow", "IMPL", "CUDA")
_libimpl.impl("erp2ico", _erp2ico)
_libimpl.impl("ico2erp_flow", _ico2erp_flow)
_libimpl.impl("warp_backward", _warp_backward)
_libimpl.impl("flow_warp_meshgrid", _flow_warp_meshgrid)
_libimpl.impl("erp_of_wraparound_", _erp_of_wraparound_)

ops = torch.ops.tangentflow
