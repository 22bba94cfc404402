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
