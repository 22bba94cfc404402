"""Global rotation helpers: oracle vs the live reference (CPU, build container) and the
CUDA path vs the oracle (GPU)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle import ref_shim, synth, tif_oracle

R_TEST = Rotation.from_euler("xyz", [0.07, -0.11, 0.03]).as_matrix()


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_against_live_reference():
    ref = ref_shim.load_reference()
    fw, sc, pr = ref["flow_warp"], ref["spherical_coordinates"], ref["projection"]
    h = 96
    _, tar = synth.synth_images(h, 4)
    for wrap in (False, True):
        assert np.array_equal(sc.rotation2erp_motion_vector((h, 2 * h), R_TEST, wrap),
                              tif_oracle.rotation2erp_motion_vector((h, 2 * h), R_TEST, wrap))
    assert np.array_equal(sc.rotate_erp_array(tar, R_TEST), tif_oracle.rotate_erp_array(tar, R_TEST))
    flow = tif_oracle.rotation2erp_motion_vector((h, 2 * h), R_TEST, True) + np.random.default_rng(0).normal(0, 0.3, (h, 2 * h, 2))
    assert np.array_equal(fw.flow2rotation_3d(flow), tif_oracle.flow2rotation_3d(flow))
    ia, ra = fw.global_rotation_warping(tar, flow, False, "3D")
    ib, rb = tif_oracle.global_rotation_warping(tar, flow, False)
    assert np.array_equal(ia, ib) and np.array_equal(ra, rb)
    for wrap in (False, True):
        assert np.array_equal(pr.flow_rotate_endpoint(flow, R_TEST.T, wrap), tif_oracle.flow_rotate_endpoint(flow, R_TEST.T, wrap))
    # the "2D" rotation type: mean longitude shift + majority-sign latitude shift of the centre band
    for shift in ((5.0, -2.0), (-7.5, 1.25), (0.0, 0.5), (2 * h - 3.0, -0.75)):      # the last one folds across the seam
        f2 = _flow_2d(h, shift)
        fa, fb = f2.copy(), f2.copy()
        ta, pa = fw.flow2rotation_2d(fa, False)
        tb, pb = tif_oracle.flow2rotation_2d(fb)
        assert (ta, pa) == (tb, pb) or (np.isnan(pa) and np.isnan(pb) and ta == tb)
        assert np.array_equal(fa, fb)                                             # the same in-place fold of u
        ia, ra = fw.global_rotation_warping(tar, f2.copy(), False, "2D")
        ib, rb = tif_oracle.global_rotation_warping(tar, f2.copy(), False, "2D")
        assert np.array_equal(ia, ib) and np.array_equal(ra, rb)
    with pytest.raises(Exception) as info:                                         # use_weight=True never worked
        fw.flow2rotation_2d(_flow_2d(h, (5.0, -2.0)), True)
    assert "axis 1 is out of bounds" in str(info.value)


def _flow_2d(h, shift, seed=3):
    rng = np.random.default_rng(seed)
    return rng.normal(0.0, 0.6, (h, 2 * h, 2)) + np.asarray(shift, dtype=np.float64)


def test_oracle_recovers_a_known_rotation():
    h = 128
    flow = tif_oracle.rotation2erp_motion_vector((h, 2 * h), R_TEST, True)
    assert np.abs(tif_oracle.flow2rotation_3d(flow) - R_TEST).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("h", [96, 512])
def test_gpu_rotation_helpers(pkg, h):
    sc, fw, pr = pkg.spherical_coordinates, pkg.flow_warp, pkg.projection
    _, tar = synth.synth_images(h, 4)
    for wrap in (False, True):
        got = sc.rotation2erp_motion_vector((h, 2 * h), R_TEST, wrap)
        want = tif_oracle.rotation2erp_motion_vector((h, 2 * h), R_TEST, wrap)
        assert got.dtype == np.float64 and np.abs(got - want).max() < 1e-9
    assert np.array_equal(sc.rotate_erp_array(tar, R_TEST), tif_oracle.rotate_erp_array(tar, R_TEST))     # uint8 exact
    flow = tif_oracle.rotation2erp_motion_vector((h, 2 * h), R_TEST, True) + np.random.default_rng(0).normal(0, 0.3, (h, 2 * h, 2))
    r_got, r_want = fw.flow2rotation_3d(flow), tif_oracle.flow2rotation_3d(flow)
    assert np.abs(r_got - r_want).max() < 1e-12
    assert np.array_equal(fw.flow2rotation_3d(flow), r_got)                                             # deterministic
    img, rot = fw.global_rotation_warping(tar, flow, forward_warp=False, rotation_type="3D")
    img_w, rot_w = tif_oracle.global_rotation_warping(tar, flow, False)
    assert img.dtype == np.uint8 and np.abs(rot - rot_w).max() < 1e-12
    assert (img != img_w).mean() < 1e-5            # the rotation differs in the last bits: at most a stray rounding
    for shift in ((5.0, -2.0), (-7.5, 1.25), (2 * h - 3.0, -0.75)):
        f2 = _flow_2d(h, shift)
        fa, fb = f2.copy(), f2.copy()
        (ta, pa), (tb, pb) = fw.flow2rotation_2d(fa, False), tif_oracle.flow2rotation_2d(fb)
        assert abs(ta - tb) < 1e-12 and abs(pa - pb) < 1e-12 and np.array_equal(fa, fb)
        assert fw.flow2rotation_2d(f2.copy(), False) == (ta, pa)                                        # deterministic
        t32, p32 = fw.flow2rotation_2d(f2.astype(np.float32), False)
        assert abs(t32 - tb) < 1e-6 and abs(p32 - pb) < 1e-6
        img2, rot2 = fw.global_rotation_warping(tar, f2.copy(), forward_warp=False, rotation_type="2D")
        img2_w, rot2_w = tif_oracle.global_rotation_warping(tar, f2.copy(), False, "2D")
        assert np.abs(rot2 - rot2_w).max() < 1e-12 and (img2 != img2_w).mean() < 1e-5
    with pytest.raises(Exception) as info:
        fw.flow2rotation_2d(_flow_2d(h, (5.0, -2.0)), True)
    assert "axis 1 is out of bounds" in str(info.value)
    for wrap in (False, True):
        got = pr.flow_rotate_endpoint(flow, R_TEST.T, wrap)
        want = tif_oracle.flow_rotate_endpoint(flow, R_TEST.T, wrap)
        err = np.abs(got - want)
        # an end point within 1e-9 px of the +-0.5 modulo edge may land on the other side (a whole image width)
        err[..., 0] = np.minimum(err[..., 0], np.abs(err[..., 0] - 2 * h))
        err[..., 1] = np.minimum(err[..., 1], np.abs(err[..., 1] - h))
        assert err.max() < 1e-8
