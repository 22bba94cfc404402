"""Stand-alone face weights projection.get_blend_weight_ico / get_blend_weight_cubemap (SURVEY.md section 8a, row a15;
the stitch kernels fuse the same weights): oracle vs the live reference (CPU, build container) and the CUDA path vs
the oracle (GPU)."""
import numpy as np
import pytest

from oracle import ref_shim, synth, tif_oracle


def _case(h, dtype, seed=11, n=5000):
    """N source / target positions around and beyond a [h, 2h] image: in range, on the last row / column exactly, and
    outside (scipy 'constant': the hard 255)."""
    rng = np.random.default_rng(seed)
    src, tar = synth.synth_images(h, seed, smooth=True)
    w = 2 * h
    x = rng.uniform(-3.0, w + 2.0, n)
    y = rng.uniform(-2.0, h + 1.5, n)
    x[:50] = w - 1.0
    y[50:100] = h - 1.0
    x[100:150] = np.floor(x[100:150])
    flow_uv = np.stack((x + rng.normal(0, 2.0, n), y + rng.normal(0, 1.5, n)), axis=1)
    flow_uv[150:200, 0] = 0.0
    if dtype != np.uint8:
        src, tar = src.astype(dtype) * dtype(1.003), tar.astype(dtype) * dtype(0.998)
    return x, y, flow_uv, src, tar


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("dtype", [np.uint8, np.float64, np.float32])
def test_oracle_against_live_reference(dtype):
    pr = ref_shim.load_reference()["projection"]
    x, y, flow_uv, src, tar = _case(48, dtype)
    want = pr.get_blend_weight_ico(x, y, "image_warp_error", flow_uv, src, tar)
    assert np.array_equal(want, tif_oracle.get_blend_weight_ico(x, y, flow_uv, src, tar))
    want = pr.get_blend_weight_cubemap(x, y, "image_warp_error", flow_uv, src, tar)
    assert np.array_equal(want, tif_oracle.get_blend_weight_cubemap(x, y, flow_uv, src, tar))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.float64, np.float32])
@pytest.mark.parametrize("h", [48, 512])
def test_gpu_blend_weights(pkg, dtype, h):
    pr = pkg.projection
    x, y, flow_uv, src, tar = _case(h, dtype, n=20000 if h > 100 else 5000)
    got = pr.get_blend_weight_ico(x, y, "image_warp_error", flow_uv, src, tar)
    want = tif_oracle.get_blend_weight_ico(x, y, flow_uv, src, tar)
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.abs(got - want).max() < 1e-12                      # weights in (0, 1]
    assert np.array_equal(got, pr.get_blend_weight_ico(x, y, "image_warp_error", flow_uv, src, tar))      # deterministic
    got = pr.get_blend_weight_cubemap(x, y, "image_warp_error", flow_uv, src, tar)
    want = tif_oracle.get_blend_weight_cubemap(x, y, flow_uv, src, tar)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    # the inner-square indicator of the cubemap stage (host geometry, no images)
    gx, gy = np.linspace(-1.2, 1.2, 41), np.linspace(1.2, -1.2, 41)
    inside = pr.get_blend_weight_cubemap(gx, gy, "straightforward")
    assert inside.sum() == ((np.abs(gx) <= 1 + 1e-7) & (np.abs(gy) <= 1 + 1e-7)).sum()
    with pytest.raises(ValueError):
        pr.get_blend_weight_ico(x, y, "cartesian_distance", flow_uv, src, tar)
