"""Drop-in replacement of the hot functions of the reference's src/flow_warp.py.

    flow_warp_meshgrid(motion_flow_u, motion_flow_v)   -> [2,H,W]
    warp_backward(image_target, of_forward)            -> image like image_target

    flow2rotation_3d(erp_flow, mask_method="center")   -> 3x3 rotation matrix          (row f2)
    global_rotation_warping(erp_image, erp_flow, forward_warp=True, rotation_type="3D")  (row f2)

numpy in -> numpy out with the reference's dtypes; torch CUDA tensors with a leading batch axis stay
on the device.  flow2rotation_2d (the "2D" rotation type) reduces its six sums on the device as well.
"""
import numpy as np
import torch

from . import _lib, _ops


def _device():
    if not torch.cuda.is_available():
        raise _lib.TifError("no CUDA device: flow_warp has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def flow_warp_meshgrid(motion_flow_u, motion_flow_v):
    """End points of the flow on the pixel grid with a single wrap (src/flow_warp.py:15-51)."""
    if isinstance(motion_flow_u, torch.Tensor):
        return _ops.ops.flow_warp_meshgrid(motion_flow_u, motion_flow_v)
    u = np.ascontiguousarray(motion_flow_u, dtype=np.float64)
    v = np.ascontiguousarray(motion_flow_v, dtype=np.float64)
    if u.shape != v.shape:
        raise ValueError("motion flow u shape %s is not equal motion flow v shape %s" % (u.shape, v.shape))
    dev = _device()
    out = _ops.ops.flow_warp_meshgrid(torch.from_numpy(u).to(dev)[None], torch.from_numpy(v).to(dev)[None])
    return out[0].cpu().numpy()


def warp_backward(image_target, of_forward):
    """Backward warp: out(y,x) = bilinear(image_target, (y+v, x+u)) (src/flow_warp.py:54-88).

    [H,W,C] images use scipy's mode='wrap' (period n-1), [H,W] images mode='constant' with 255;
    uint8 images are rounded half-up; float32 / float64 images are accumulated in float64 in scipy's order and
    stored in their own dtype, so the result equals the reference's bit for bit.
    """
    if isinstance(image_target, torch.Tensor):
        mode = _lib.WARP_WRAP
        return _ops.ops.warp_backward(image_target, of_forward, mode)
    image_target = np.asarray(image_target)
    flow = np.ascontiguousarray(of_forward)
    if flow.dtype not in (np.float32, np.float64):
        flow = flow.astype(np.float64)
    if image_target.ndim == 3:
        mode, img = _lib.WARP_WRAP, image_target
    elif image_target.ndim == 2:
        mode, img = _lib.WARP_CONSTANT255, image_target[:, :, None]
    else:
        raise ValueError("The image shape is %s, do not support." % (image_target.shape,))
    dev = _device()
    if img.dtype in (np.uint8, np.float32, np.float64):
        img_dev = torch.from_numpy(np.ascontiguousarray(img)).to(dev)[None]
    else:           # other dtypes: scipy computes in float64 and casts back to the image's dtype
        img_dev = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float64)).to(dev)[None]
    out = _ops.ops.warp_backward(img_dev, torch.from_numpy(flow).to(dev)[None], mode)[0].cpu().numpy()
    out = out.astype(image_target.dtype, copy=False)
    return out if image_target.ndim == 3 else out[:, :, 0]


def rotation_from_covariance(h_mat):
    """R = V U^T from the SVD of the 3x3 cross covariance (src/pointcloud_utils.py:31-32); 9 numbers, on the host."""
    u, _, vt = np.linalg.svd(np.asarray(h_mat, dtype=np.float64).reshape(3, 3))
    return np.dot(vt.T, u.T)


def flow2rotation_3d(erp_flow, mask_method="center"):
    """Rotation between two ERP images from their flow (src/flow_warp.py:91-133): cross covariance of the pixel
    directions and the flow end-point directions over the middle half of the rows (device reduction), then SVD."""
    if mask_method != "center":
        raise NotImplementedError("flow2rotation_3d: only mask_method='center' is on the accelerated path")
    if isinstance(erp_flow, torch.Tensor):
        h = _ops.flow_cross_covariance(erp_flow if erp_flow.dim() == 4 else erp_flow[None])
        mats = np.stack([rotation_from_covariance(m) for m in h.cpu().numpy()])
        return mats if erp_flow.dim() == 4 else mats[0]
    flow = np.ascontiguousarray(erp_flow)
    if flow.dtype not in (np.float32, np.float64):
        flow = flow.astype(np.float64)
    h = _ops.flow_cross_covariance(torch.from_numpy(flow).to(_device())[None])
    return rotation_from_covariance(h[0].cpu().numpy())


def flow2rotation_2d(erp_flow, use_weight=True):
    """(theta shift, phi shift) in radians between two ERP images from their flow (src/flow_warp.py:136-187): the
    mean longitude shift, then the mean latitude shift of the majority sign inside the centre band of columns the
    longitude shift spans.  Like the reference it first folds the flow with erp_of_wraparound (which rewrites the
    caller's u plane).  `use_weight=True` fails in the reference (numpy AxisError at :183: a 1-D array averaged
    over axis 1), so it raises here too; `global_rotation_warping` passes False."""
    from . import flow_postproc
    if use_weight:
        raise np.exceptions.AxisError(1, 1) if hasattr(np, "exceptions") else ValueError("axis 1 is out of bounds for array of dimension 1")
    flow = flow_postproc.erp_of_wraparound(erp_flow)
    h, w = flow.shape[0], flow.shape[1]
    if flow.dtype not in (np.float32, np.float64):
        flow = flow.astype(np.float64)
    dev_flow = torch.from_numpy(np.ascontiguousarray(flow)).to(_device())[None]
    sum_u = float(_ops.flow_rotation2d_stats(dev_flow, 0, 0)[0, 0].item())
    theta_delta = 2.0 * np.pi * (sum_u / (h * w) / w)
    delta = theta_delta / (2.0 * np.pi)
    col_start, col_end = int(w * (0.5 - delta)), int(w * (0.5 + delta))
    if delta < 0:
        col_start, col_end = col_end, col_start
    st = _ops.flow_rotation2d_stats(dev_flow, col_start, col_end)[0].cpu().numpy()
    total, count = (st[4], st[5]) if st[1] < 0 else (st[2], st[3])
    with np.errstate(invalid="ignore", divide="ignore"):
        phi_delta = -np.pi * (np.float64(total) / np.float64(count)) / h      # nan on an empty band, like np.mean([])
    return np.float64(theta_delta), np.float64(phi_delta)


def global_rotation_warping(erp_image, erp_flow, forward_warp=True, rotation_type="3D"):
    """Rotate the ERP image by the global rotation of the flow (src/flow_warp.py:190-229).  Returns the rotated
    image and the rotation matrix."""
    from . import spherical_coordinates as sc
    if rotation_type == "2D":
        theta_delta, phi_delta = flow2rotation_2d(erp_flow, False)
        if not forward_warp:
            theta_delta, phi_delta = -theta_delta, -phi_delta
        rotation_mat = sc.rot_sph2mat(theta_delta, phi_delta, False)
    elif rotation_type == "3D":
        rotation_mat = flow2rotation_3d(erp_flow)
        if not forward_warp:
            rotation_mat = rotation_mat.T
    else:
        raise UnboundLocalError("rotation_mat: the reference logs 'Do not suport rotation type %s' and then fails" % (rotation_type,))
    erp_image_rot = sc.rotate_erp_array(erp_image, rotation_mat=rotation_mat)
    if erp_image.dtype == np.uint8:
        erp_image_rot = erp_image_rot.astype(np.uint8)
    return erp_image_rot, rotation_mat
