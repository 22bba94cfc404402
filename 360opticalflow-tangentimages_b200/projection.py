"""Drop-in replacement of `flow_rotate_endpoint` of the reference's src/projection.py (row f2).

The blend-weight functions of that module (get_blend_weight_ico / get_blend_weight_cubemap) have no counterpart
here: they are fused into k_lift / k_gather (csrc/tif_exec.cu).
"""
import numpy as np
import torch

from . import _lib, _ops
from . import spherical_coordinates as sc


def flow_rotate_endpoint(optical_flow, rotation, wraparound=False):
    """Add a rotation to the end points of an ERP flow (src/projection.py:120-159).  `rotation` is a 3x3 matrix or
    a (theta, phi) pair in degrees (as rot_sph2mat's default)."""
    if not torch.cuda.is_available():
        raise _lib.TifError("no CUDA device: flow_rotate_endpoint has no CPU fallback")
    if isinstance(rotation, (list, tuple)):
        rotation = sc.rot_sph2mat(rotation[0], rotation[1])
    elif not isinstance(rotation, np.ndarray):
        raise TypeError("Do not support rotation data type %s." % type(rotation))
    dev = torch.device("cuda", torch.cuda.current_device())
    as_numpy = not isinstance(optical_flow, torch.Tensor)
    if as_numpy:
        flow = np.ascontiguousarray(optical_flow)
        if flow.dtype not in (np.float32, np.float64):
            flow = flow.astype(np.float64)
        flow_dev = torch.from_numpy(flow).to(dev)[None]
    else:
        flow_dev = optical_flow if optical_flow.dim() == 4 else optical_flow[None]
    h, w = flow_dev.shape[1:3]
    field = _ops.rotation_motion_vector(rotation, h, w, True, flow_dev.device)
    out = _ops.flow_rotate_endpoint(flow_dev, field, wraparound)
    if as_numpy:
        return out[0].cpu().numpy()
    return out if optical_flow.dim() == 4 else out[0]


def _warp_error_weights(face_x_src, face_y_src, flow_uv, image_erp_src, image_erp_tar, kind):
    if not torch.cuda.is_available():
        raise _lib.TifError("no CUDA device: projection has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    src, tar = np.asarray(image_erp_src), np.asarray(image_erp_tar)
    if src.dtype != tar.dtype or src.dtype not in (np.uint8, np.float32, np.float64):
        src, tar = src.astype(np.float64), tar.astype(np.float64)
    if src.ndim != 3 or src.shape != tar.shape:
        raise ValueError("get_blend_weight: image_erp_src / image_erp_tar must be [H,W,C] arrays of one shape")
    flow_uv = np.asarray(flow_uv, dtype=np.float64)
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    out = _ops.blend_weight_points(torch.from_numpy(np.ascontiguousarray(src)).to(dev), torch.from_numpy(np.ascontiguousarray(tar)).to(dev),
                                   to_dev(face_x_src), to_dev(face_y_src), to_dev(flow_uv[:, 0]), to_dev(flow_uv[:, 1]), kind)
    return out.cpu().numpy()


def get_blend_weight_ico(face_x_src_gnomonic, face_y_src_gnomonic, weight_type, flow_uv=None, image_erp_src=None,
                         image_erp_tar=None, gnomonic_bounding_box=None):
    """Photometric face weights exp(-mean_c|tar(flow end point) - src(pixel)| / mean over the face)
    (src/projection.py:15-61; despite their names the first two arguments are ERP pixel coordinates, like
    `flow_uv`).  ico2erp_flow fuses the same computation; this is the stand-alone form."""
    if weight_type != "image_warp_error":
        raise ValueError("the weight method %s do not exist." % (weight_type,))       # the reference logs this and returns zeros
    return _warp_error_weights(face_x_src_gnomonic, face_y_src_gnomonic, flow_uv, image_erp_src, image_erp_tar, 0)


def get_blend_weight_cubemap(face_x_src_gnomonic, face_y_src_gnomonic, weight_type, flow_uv=None, image_erp_src=None,
                             image_erp_tar=None, gnomonic_bounding_box=None):
    """Cubemap face weights (src/projection.py:64-117): "straightforward" = 1 inside the face square (gnomonic
    coordinates, inclusive, eps 1e-7), "image_warp_error" = 0.95 / ||tar - src||_2 (1 where the error is 0)."""
    if weight_type == "straightforward":
        from . import polygon
        if gnomonic_bounding_box is None:
            gnomonic_bounding_box = np.array([[-1, 1], [1, 1], [1, -1], [-1, -1]])
        inside = polygon.inside_polygon_2d(np.stack((face_x_src_gnomonic, face_y_src_gnomonic), axis=1), gnomonic_bounding_box,
                                           on_line=True, eps=1e-7)
        weight_map = np.zeros(np.shape(face_x_src_gnomonic)[0], dtype=np.float64)
        weight_map[inside] = 1.0
        return weight_map
    if weight_type != "image_warp_error":
        raise ValueError("the weight method %s do not exist." % (weight_type,))
    return _warp_error_weights(face_x_src_gnomonic, face_y_src_gnomonic, flow_uv, image_erp_src, image_erp_tar, 1)
