"""Drop-in replacement of the metric functions of the reference's src/flow_evaluate.py (row f3).

    available_pixel(flow, of_mask=None, unknown_value=1e7)
    EPE / EPE_mat / RMSE / RMSE_mat / AAE / AAE_mat (of_ground_truth, of_evaluation, spherical=False, of_mask=None)

Per-pixel matrices and their sums are computed on the device (csrc/tif_metrics.cu).  Quirks of the reference
that are kept: available_pixel's operator precedence, the *_mat functions zero the unusable pixels of BOTH input
arrays in place, AAE ignores its `spherical` and `of_mask` arguments, the spherical AAE_mat branch does not exist
(the reference calls a missing function there).
"""
import numpy as np
import torch

from . import _lib, _ops

UNKNOWN_FLOW_THRESH = 1e7


def _device():
    if not torch.cuda.is_available():
        raise _lib.TifError("no CUDA device: flow_evaluate has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _error_mats(of_ground_truth, of_evaluation, spherical, of_mask, want_aae):
    dev = _device()
    lib = _lib.load()
    gt = np.ascontiguousarray(of_ground_truth, dtype=np.float64)
    ev = np.ascontiguousarray(of_evaluation, dtype=np.float64)
    if gt.shape != ev.shape or gt.ndim != 3 or gt.shape[2] != 2:
        raise ValueError("flows must both be [H,W,2]")
    h, w = gt.shape[:2]
    gt_d, ev_d = torch.from_numpy(gt).to(dev), torch.from_numpy(ev).to(dev)
    mask_d = None
    if of_mask is not None:
        mask_d = torch.from_numpy(np.ascontiguousarray(np.asarray(of_mask) != 0).astype(np.uint8)).to(dev)
    epe = torch.empty((h, w), dtype=torch.float64, device=dev)
    aae = torch.empty((h, w), dtype=torch.float64, device=dev) if want_aae else None
    valid = torch.empty((h, w), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.tif_flow_error_mats(gt_d.data_ptr(), ev_d.data_ptr(), _lib.DTYPE_F64,
                                           mask_d.data_ptr() if mask_d is not None else None, 1, h, w, int(bool(spherical)),
                                           epe.data_ptr(), aae.data_ptr() if aae is not None else None, valid.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "tif_flow_error_mats")
    ok = valid.cpu().numpy().astype(bool)
    # like the reference, the unusable pixels of both caller arrays are zeroed (its *_mat functions write through views)
    for arr in (of_ground_truth, of_evaluation):
        if isinstance(arr, np.ndarray) and arr.flags.writeable:
            arr[~ok] = 0
    return epe, aae, valid, ok


def _sum(x, power=1):
    lib = _lib.load()
    x = x.contiguous().view(1, -1)
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    scratch = torch.empty(int(lib.tif_reduce_f64_scratch_bytes(1)), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.tif_reduce_f64(x.data_ptr(), 1, x.shape[1], power, out.data_ptr(), scratch.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream), "tif_reduce_f64")
    return float(out.cpu()[0])


def available_pixel(flow, of_mask=None, unknown_value=UNKNOWN_FLOW_THRESH):
    """src/flow_evaluate.py:23-54 (boolean [H,W]); evaluated with the error kernel against a zero flow."""
    if unknown_value != UNKNOWN_FLOW_THRESH:
        raise NotImplementedError("available_pixel: only the default unknown_value is on the accelerated path")
    flow = np.array(flow, dtype=np.float64, copy=True)
    _, _, _, ok = _error_mats(flow, np.zeros_like(flow), False, of_mask, False)
    return ok


def EPE_mat(of_ground_truth, of_evaluation, spherical=False, of_mask=None):
    """src/flow_evaluate.py:70-104 -> (epe float64 [H,W], available bool [H,W])."""
    epe, _, _, ok = _error_mats(of_ground_truth, of_evaluation, spherical, of_mask, False)
    return epe.cpu().numpy(), ok


def EPE(of_ground_truth, of_evaluation, spherical=False, of_mask=None):
    """Mean end-point error over the available pixels (src/flow_evaluate.py:57-67)."""
    epe, _, valid, ok = _error_mats(of_ground_truth, of_evaluation, spherical, of_mask, False)
    return _sum(epe) / float(ok.sum())


def RMSE_mat(of_ground_truth, of_evaluation, spherical=False, of_mask=None):
    """src/flow_evaluate.py:118-153 (same matrix as EPE_mat)."""
    return EPE_mat(of_ground_truth, of_evaluation, spherical, of_mask)


def RMSE(of_ground_truth, of_evaluation, spherical=False, of_mask=None):
    """src/flow_evaluate.py:107-115."""
    epe, _, valid, ok = _error_mats(of_ground_truth, of_evaluation, spherical, of_mask, False)
    return float(np.sqrt(_sum(epe, 2) / float(ok.sum())))


def AAE_mat(of_ground_truth, of_evaluation, spherical=False, of_mask=None):
    """src/flow_evaluate.py:167-217, planar branch (the spherical branch of the reference calls a function that
    does not exist)."""
    if spherical:
        raise AttributeError("module 'spherical_coordinates' has no attribute 'get_angle'")     # what the reference raises
    _, aae, _, ok = _error_mats(of_ground_truth, of_evaluation, False, of_mask, True)
    return aae.cpu().numpy(), ok


def AAE(of_ground_truth, of_evaluation, spherical=False, of_mask=None):
    """src/flow_evaluate.py:156-164: always planar and unmasked, whatever is passed."""
    _, aae, _, ok = _error_mats(of_ground_truth, of_evaluation, False, None, True)
    return _sum(aae) / float(ok.sum())
