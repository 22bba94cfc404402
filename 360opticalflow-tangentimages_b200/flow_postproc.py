"""Drop-in replacement of the reference's src/flow_postproc.py.

    erp_pixles_modulo(x_arrray, y_array, image_width, image_height)
    erp_of_wraparound(erp_flow, of_u_threshold=None)
"""
import numpy as np
import torch

from . import _lib, _ops


def erp_pixles_modulo(x_arrray, y_array, image_width, image_height):
    """Wrap pixel coordinates into the ERP image (src/flow_postproc.py:9-14).  On the hot path this is
    the integer `wrapped_multiplicity` logic of k_plan_stage2; this host form serves callers' small arrays."""
    return np.remainder(x_arrray + 0.5, image_width) - 0.5, np.remainder(y_array + 0.5, image_height) - 0.5


def erp_of_wraparound(erp_flow, of_u_threshold=None):
    """Fold |u| > threshold back across the 360 degree seam (src/flow_postproc.py:17-44).

    Like the reference it rewrites the u plane of its (numpy) argument and returns a new stacked array;
    the last column is never touched.  A torch CUDA tensor [B,H,W,2] is updated in place and returned.
    """
    if isinstance(erp_flow, torch.Tensor):
        thr = erp_flow.shape[2] / 2.0 if of_u_threshold is None else float(of_u_threshold)
        return _ops.ops.erp_of_wraparound_(erp_flow, thr)
    if not torch.cuda.is_available():
        raise _lib.TifError("no CUDA device: flow_postproc has no CPU fallback")
    width = np.shape(erp_flow)[1]
    thr = width / 2.0 if of_u_threshold is None else float(of_u_threshold)
    dev = torch.device("cuda", torch.cuda.current_device())
    dtype = erp_flow.dtype if erp_flow.dtype in (np.float32, np.float64) else np.float64
    work = torch.from_numpy(np.ascontiguousarray(erp_flow, dtype=dtype)).to(dev)[None]
    out = _ops.ops.erp_of_wraparound_(work, thr)[0].cpu().numpy()
    erp_flow[:, :, 0] = out[:, :, 0]          # the reference mutates the caller's u plane through a view
    return np.stack((out[:, :, 0], erp_flow[:, :, 1]), axis=2)
