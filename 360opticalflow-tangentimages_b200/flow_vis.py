"""Drop-in replacement of the reference's src/flow_vis.py (row f4): Middlebury colour-wheel rendering of a flow field.

    make_colorwheel()                                      -> [55,3] float64
    flow_uv_to_colors(u, v, convert_to_bgr=False, sph_of=False) -> uint8 [H,W,3]
    flow_to_color(flow_uv, clip_flow=None, convert_to_bgr=False, min_ratio=0.0, max_ratio=1.0, add_bar=False, sph_of=False)
    create_colorwheel_bar(image_size)

The per-pixel mapping (radius, angle, wheel interpolation, floor(255 c)) runs on the device in float64 in the
reference's operation order (csrc/tif_vis.cu); the 55-row wheel and the quantile clip range (image_evaluate.get_min_max,
an np.partition) stay on the host.  float64 flows -- what ico2erp_flow returns and main.py renders -- give the
reference's bytes; float32 flows are evaluated in float64 here while numpy keeps them in float32, which may move a
value by one grey level where 255 c falls within 1e-5 of an integer.  `add_bar=True` draws through matplotlib in the
reference (not installed here) and is not provided.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def make_colorwheel():
    """The 55-colour wheel of Baker et al. (src/flow_vis.py:19-66)."""
    segments = (15, 6, 4, 11, 13, 6)            # RY, YG, GC, CB, BM, MR
    wheel = np.zeros((sum(segments), 3))
    # (channel that ramps, direction, channel held at 255) per segment
    plan = ((1, +1, 0), (0, -1, 1), (2, +1, 1), (1, -1, 2), (0, +1, 2), (2, -1, 0))
    col = 0
    for n, (ramp, sign, full) in zip(segments, plan):
        steps = np.floor(255 * np.arange(0, n) / n)
        wheel[col:col + n, full] = 255
        wheel[col:col + n, ramp] = steps if sign > 0 else 255 - steps
        col += n
    return wheel


def get_min_max(data, min_ratio=0.0, max_ratio=1.0):
    """Quantile clip range of the reference's image_evaluate.get_min_max (src/image_evaluate.py:8-32)."""
    if min_ratio != 0.0 or max_ratio != 1.0:
        flat = np.asarray(data).flatten()
        lo_idx, hi_idx = int(flat.size * min_ratio), int((flat.size - 1) * max_ratio)
        return np.partition(flat, lo_idx)[lo_idx], np.partition(flat, hi_idx)[hi_idx]
    return np.amin(data), np.amax(data)


def _render(flow_uv, clip, norm, convert_to_bgr):
    if not torch.cuda.is_available():
        raise _lib.TifError("no CUDA device: flow_vis has no CPU fallback")
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    flow = np.ascontiguousarray(flow_uv)
    if flow.dtype not in (np.float32, np.float64):
        flow = flow.astype(np.float64)
    h, w = flow.shape[:2]
    d_flow = torch.from_numpy(flow).to(dev)
    dtype = _lib.DTYPE_F32 if flow.dtype == np.float32 else _lib.DTYPE_F64
    use_clip, lo, hi = (1, float(clip[0]), float(clip[1])) if clip is not None else (0, 0.0, 0.0)
    stream = torch.cuda.current_stream().cuda_stream
    with torch.cuda.device(dev):
        if norm is None:            # normalise by the largest radius (src/flow_vis.py:157-162)
            d_max = torch.zeros(1, dtype=torch.float64, device=dev)
            _lib.check(lib.tif_flow_rad_max(d_flow.data_ptr(), dtype, h * w, use_clip, lo, hi, d_max.data_ptr(), stream), "tif_flow_rad_max")
            rad_max = float(d_max.item())
            if flow.dtype == np.float32:
                rad_max = float(np.float32(rad_max))
            norm = rad_max + 1e-5
        wheel = torch.from_numpy(make_colorwheel()).to(dev)
        out = torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
        _lib.check(lib.tif_flow_to_color_u8(d_flow.data_ptr(), dtype, h * w, use_clip, lo, hi, float(norm), wheel.data_ptr(),
                                            int(bool(convert_to_bgr)), out.data_ptr(), stream), "tif_flow_to_color_u8")
    return out.cpu().numpy()


def flow_uv_to_colors(u, v, convert_to_bgr=False, sph_of=False):
    """Colour-wheel image of (possibly clipped, already normalised) flow components (src/flow_vis.py:77-121)."""
    if sph_of:
        raise NotImplementedError("the reference raises NotImplemented for sph_of=True (src/flow_vis.py:105)")
    return _render(np.stack((np.asarray(u), np.asarray(v)), axis=-1), None, 0.0, convert_to_bgr)


def create_colorwheel_bar(image_size):
    """The wheel itself as an image (src/flow_vis.py:68-74)."""
    axis = np.linspace(-1.0, 1.0, num=image_size, endpoint=False)
    xv, yv = np.meshgrid(axis, axis)
    return flow_uv_to_colors(xv, yv)


def flow_to_color(flow_uv, clip_flow=None, convert_to_bgr=False, min_ratio=0.0, max_ratio=1.0, add_bar=False, sph_of=False):
    """Flow field [H,W,2] -> uint8 [H,W,3] (src/flow_vis.py:124-201)."""
    if add_bar:
        raise NotImplementedError("flow_to_color(add_bar=True) draws through matplotlib in the reference; not provided")
    if sph_of:
        raise NotImplementedError("the reference raises NotImplemented for sph_of=True (src/flow_vis.py:105)")
    flow_uv = np.asarray(flow_uv)
    if min_ratio != 0 and max_ratio != 1.0:          # the reference's `and` (src/flow_vis.py:140)
        clip_flow = get_min_max(flow_uv, min_ratio, max_ratio)
    assert flow_uv.ndim == 3, 'input flow must have three dimensions'
    assert flow_uv.shape[2] == 2, 'input flow must have shape [H,W,2]'
    return _render(flow_uv, clip_flow, None, convert_to_bgr)
