"""Drop-in replacement of the metric functions of the reference's src/flow_evaluate.py (row f3).

    available_pixel(flow, of_mask=None, unknown_value=1e7)
    EPE / EPE_mat / RMSE / RMSE_mat / AAE / AAE_mat (of_ground_truth, of_evaluation, spherical=False, of_mask=None)

Per-pixel matrices and their sums are computed on the device (csrc/tif_metrics.cu), in float64 whatever the dtype of
the flows: float32 flows (e.g. read from .flo files) are upcast first, where the reference sums them in float32, so
the means agree with the reference to ~1e-7 relative for float32 input and to 1e-11 for float64.  Quirks of the reference
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
        # the reference combines `index_valid & of_mask` (src/flow_evaluate.py:48): a bitwise AND with a boolean, so an
        # integer mask counts through its lowest bit only (2 or 254 mark a pixel unusable), and a float mask is a TypeError
        m = np.asarray(of_mask)
        if m.dtype == np.bool_:
            m = m.astype(np.uint8)
        elif np.issubdtype(m.dtype, np.integer):
            m = (m & 1).astype(np.uint8)
        else:
            raise TypeError("ufunc 'bitwise_and' not supported for the input types (bool & %s)" % m.dtype)
        mask_d = torch.from_numpy(np.ascontiguousarray(m)).to(dev)
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
    with np.errstate(invalid="ignore", divide="ignore"):
        return float(np.float64(_sum(epe)) / np.float64(ok.sum()))        # nan on an empty set, like np.mean([])


def RMSE_mat(of_ground_truth, of_evaluation, spherical=False, of_mask=None):
    """src/flow_evaluate.py:118-153 (same matrix as EPE_mat)."""
    return EPE_mat(of_ground_truth, of_evaluation, spherical, of_mask)


def RMSE(of_ground_truth, of_evaluation, spherical=False, of_mask=None):
    """src/flow_evaluate.py:107-115."""
    epe, _, valid, ok = _error_mats(of_ground_truth, of_evaluation, spherical, of_mask, False)
    with np.errstate(invalid="ignore", divide="ignore"):
        return float(np.sqrt(np.float64(_sum(epe, 2)) / np.float64(ok.sum())))


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
    with np.errstate(invalid="ignore", divide="ignore"):
        return float(np.float64(_sum(aae)) / np.float64(ok.sum()))


def opticalflow_metric(flo_gt, flo_eva, erp_flo_error_filename_prefix=None, flo_mask=None, min_ratio=0.1, max_ratio=0.9,
                       verbose=True):
    """AAE, EPE, RMS and their "spherical" variants in the reference's order (src/flow_evaluate.py:220-280); like the
    reference every metric zeroes the unusable pixels of both flows in place, so the later ones see what the earlier
    ones left.  With a file name prefix the reference also writes colour-mapped error images through its own
    `image_io` module (used as it is when it is importable) -- and then fails in `AAE_mat(spherical=True)`, which
    calls a function that does not exist; so does this."""
    aae = AAE(flo_gt, flo_eva, spherical=False, of_mask=flo_mask)
    epe = EPE(flo_gt, flo_eva, spherical=False, of_mask=flo_mask)
    rms = RMSE(flo_gt, flo_eva, spherical=False, of_mask=flo_mask)
    aae_sph = AAE(flo_gt, flo_eva, spherical=True, of_mask=flo_mask)
    epe_sph = EPE(flo_gt, flo_eva, spherical=True, of_mask=flo_mask)
    rms_sph = RMSE(flo_gt, flo_eva, spherical=True, of_mask=flo_mask)
    if erp_flo_error_filename_prefix is not None:
        import image_io            # the reference's module (reference/src on sys.path); not part of this package
        for name, fn, spherical in (("_aae_mat.jpg", AAE_mat, False), ("_aae_mat_sph.jpg", AAE_mat, True),
                                    ("_epe_mat.jpg", EPE_mat, False), ("_epe_mat_sph.jpg", EPE_mat, True),
                                    ("_rms_mat.jpg", RMSE_mat, False), ("_rms_mat_sph.jpg", RMSE_mat, True)):
            mat, _ = fn(flo_gt, flo_eva, spherical, flo_mask)
            image_io.image_save(image_io.visual_data(mat, min_ratio, max_ratio), erp_flo_error_filename_prefix + name)
    return aae, epe, rms, aae_sph, epe_sph, rms_sph


def opticalflow_metric_folder(of_dir, of_gt_dir, mask_filename_exp=None, result_csv_filename=None, visual_of_error=False,
                              of_wraparound=True, skip_list=[]):
    """The six metrics of every `.flo` file of a folder against the ground truth of the same name, one CSV row each
    (src/flow_evaluate.py:283-350)."""
    import csv
    import os
    from . import flow_io, flow_postproc
    path = of_dir + ("result.csv" if result_csv_filename is None else result_csv_filename)
    with open(path, mode="w", newline="", encoding="utf-8") as result_file:
        writer = csv.writer(result_file, delimiter=",", quotechar='"', quoting=csv.QUOTE_MINIMAL)
        writer.writerow(["flo_filename", "AAE", "EPE", "RMS", "SAAE", "SEPE", "SRMS"])
        for filename in sorted(os.listdir(of_dir)):
            if not filename.endswith(".flo"):
                continue
            if filename in skip_list:
                print("skip {} in {}".format(filename, skip_list))
                continue
            flow = flow_io.read_flow_flo(of_dir + filename)
            flow_gt = flow_io.read_flow_flo(of_gt_dir + filename)
            if flow.shape != flow_gt.shape:        # the reference calls flow_postproc.flow_resize, which does not exist
                raise AttributeError("module 'flow_postproc' has no attribute 'flow_resize'")
            if of_wraparound:
                flow = flow_postproc.erp_of_wraparound(flow)
                flow_gt = flow_postproc.erp_of_wraparound(flow_gt)
            mask = None
            if mask_filename_exp is not None:
                import image_io        # the reference's module
                mask = image_io.image_read(of_gt_dir + mask_filename_exp.format(int(filename[0:4])))
            prefix = of_dir + filename if visual_of_error else None
            writer.writerow([filename, *opticalflow_metric(flow_gt, flow, prefix, mask, 0.1, 0.9, True)])
            result_file.flush()
