"""Row f4: colour-wheel visualisation (src/flow_vis.py:19-201).  CPU: the oracle restatement against the live reference;
GPU: the device rendering against the oracle -- identical bytes for float64 flows."""
import importlib

import numpy as np
import pytest

from oracle import ref_shim, tif_oracle


def _flow(h, w, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    f = rng.normal(0.0, 7.0, (h, w, 2))
    f[0, 0] = 0.0                       # zero flow: atan2(-0, -0)
    f[1, :, 0] = 0.0
    f[:, 2, 1] = 0.0
    f[3, 3] = (250.0, -130.0)           # an outlier that the quantile clip removes
    return f.astype(dtype)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_flow_vis_against_live_reference():
    ref_shim.load_reference()
    fv = importlib.import_module("flow_vis")
    assert np.array_equal(fv.make_colorwheel(), tif_oracle.make_colorwheel())
    for seed, dtype in ((1, np.float64), (2, np.float32)):
        f = _flow(48, 96, seed, dtype)
        assert np.array_equal(fv.flow_to_color(f.copy()), tif_oracle.flow_to_color(f.copy()))
        assert np.array_equal(fv.flow_to_color(f.copy(), min_ratio=0.2, max_ratio=0.8),
                              tif_oracle.flow_to_color(f.copy(), min_ratio=0.2, max_ratio=0.8))
        assert np.array_equal(fv.flow_to_color(f.copy(), clip_flow=[-5, 5], convert_to_bgr=True),
                              tif_oracle.flow_to_color(f.copy(), clip_flow=[-5, 5], convert_to_bgr=True))


def test_product_colorwheel_matches_oracle(pkg):
    assert np.array_equal(pkg.flow_vis.make_colorwheel(), tif_oracle.make_colorwheel())
    lo, hi = pkg.flow_vis.get_min_max(np.arange(100.0).reshape(10, 10), 0.2, 0.8)
    assert (lo, hi) == (20.0, 79.0)


@pytest.mark.gpu
def test_flow_to_color_on_device(pkg):
    vis = pkg.flow_vis
    f = _flow(200, 400, 3)
    assert np.array_equal(vis.flow_to_color(f.copy()), tif_oracle.flow_to_color(f.copy()))
    assert np.array_equal(vis.flow_to_color(f.copy(), min_ratio=0.2, max_ratio=0.8), tif_oracle.flow_to_color(f.copy(), min_ratio=0.2, max_ratio=0.8))
    assert np.array_equal(vis.flow_to_color(f.copy(), clip_flow=[-5, 5], convert_to_bgr=True),
                          tif_oracle.flow_to_color(f.copy(), clip_flow=[-5, 5], convert_to_bgr=True))
    wheel = vis.create_colorwheel_bar(150)
    axis = np.linspace(-1.0, 1.0, num=150, endpoint=False)
    xv, yv = np.meshgrid(axis, axis)
    assert np.array_equal(wheel, tif_oracle.flow_uv_to_colors(xv, yv))
    # float32 flows: evaluated in float64 on the device, in float32 by numpy -- single grey levels may differ where
    # 255 c lies within 1e-5 of an integer
    f32 = _flow(200, 400, 4, np.float32)
    got, want = vis.flow_to_color(f32.copy()), tif_oracle.flow_to_color(f32.copy())
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3
    with pytest.raises(NotImplementedError):
        vis.flow_to_color(f, add_bar=True)
