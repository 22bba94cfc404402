"""The oracle against the committed outputs of the live reference (tests/golden/, made by
oracle/make_golden.py) and, where /root/reference is mounted, against the reference itself."""
import hashlib

import numpy as np
import pytest
from scipy import ndimage

from oracle import ref_shim, synth, tif_oracle

SMALL = ["small_h64_noise", "small_h80_smooth_seam", "small_h48_pad01"]
SUMMARY = ["summary_h640_smooth_seam", "summary_h1024_noise", "summary_h1920_noise", "summary_h2048_smooth"]
FLOAT_TOL = 1e-9     # numpy's arcsin/arctan2 differ in the last ulp between CPU dispatch targets


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs(g):
    src, tar = synth.synth_images(int(g["erp_h"]), int(g["seed"]), bool(g["smooth"]))
    flows = synth.synth_flows(int(g["tangent_w"]), 20, float(g["u_offset"]))
    assert sha(src) == str(g["sha_src"]) and sha(tar) == str(g["sha_tar"]) and sha(np.stack(flows)) == str(g["sha_flows"])
    return src, tar, flows


def oracle_membership(flows, erp_h, pad):
    _, dbg = tif_oracle.ico2erp_flow(flows, erp_h, pad, None, None, False, "straightforward", return_debug=True)
    return dbg["faces"]


def test_face_table_matches_reference(golden):
    g = golden("face_table")
    for pi_, pad in enumerate(g["pads"]):
        for f in range(20):
            par = tif_oracle.icosahedron_face(f, float(pad))
            assert np.array_equal([par["theta_0"], par["phi_0"]], g["tangent_point"][pi_, f])
            assert np.array_equal(par["tri"], g["tri"][pi_, f])
            assert np.array_equal(np.array(par["area"]), g["area"][pi_, f])
            for hi, h in enumerate(g["bbox_heights"]):
                assert tuple(g["bbox"][pi_, hi, f]) == tif_oracle.face_erp_bbox(par, int(h))


@pytest.mark.parametrize("mode", ["wrap", "constant", "nearest"])
def test_map_coordinates_restatement_matches_scipy(mode):
    rng = np.random.default_rng(1)
    for img in (rng.integers(0, 256, (37, 53), dtype=np.uint8), rng.normal(size=(37, 53))):
        ys = rng.uniform(-80, 120, 50000)
        xs = rng.uniform(-120, 180, 50000)
        ys[:2000] = rng.integers(-40, 80, 2000)
        xs[:2000] = rng.integers(-60, 110, 2000)
        ys[3000:3100], xs[3100:3200], ys[3200:3300], xs[3300:3400] = 36.0, 52.0, -0.5, 52.5
        if mode == "nearest":     # the reference only uses 'nearest' at integer coordinates (an exact gather)
            ys, xs = np.round(ys), np.round(xs)
        want = ndimage.map_coordinates(img, [ys, xs], order=1, mode=mode, cval=255)
        got = tif_oracle.map_coordinates_order1(img, ys, xs, mode, 255)
        assert got.dtype == want.dtype and np.array_equal(got, want)


@pytest.mark.parametrize("name", SMALL)
def test_small_cases(golden, name):
    g = golden(name)
    src, tar, flows = inputs(g)
    erp_h, wt, pad = int(g["erp_h"]), int(g["tangent_w"]), float(g["pad"])
    full = tif_oracle.erp2ico_image(src, wt, pad, True)
    assert all(t.dtype == np.int64 for t in full)
    assert np.array_equal(np.stack(full), g["ico_full"].astype(np.int64))
    assert np.array_equal(np.stack(tif_oracle.erp2ico_image(src, wt, pad, False)), g["ico_part"].astype(np.int64))
    for blend in ("straightforward", "normwarp"):
        for wrap in (False, True):
            f = tif_oracle.ico2erp_flow(flows, erp_h, pad, src, tar, wrap, blend)
            assert np.abs(f - g["flow_%s_%d" % (blend[:2], int(wrap))]).max() < FLOAT_TOL
    faces = oracle_membership(flows, erp_h, pad)
    got = np.concatenate([np.stack((np.full(d["rows"].shape, i), d["rows"], d["cols"], d["ix"], d["iy"], d["mult"]), 1)
                          for i, d in enumerate(faces)])
    want = np.stack((g["mem_face"], g["mem_row"], g["mem_col"], g["mem_ix"], g["mem_iy"], g["mem_mult"]), 1)
    assert np.array_equal(got, want)
    assert np.array_equal([d["candidates"] for d in faces], g["candidates"])
    assert np.array_equal([d["bbox"] for d in faces], g["bbox"])
    nw = g["flow_no_1"]
    assert np.array_equal(tif_oracle.warp_backward(tar, nw), g["warp_u8"])
    assert np.array_equal(tif_oracle.warp_backward(tar[:, :, 0].copy(), nw), g["warp_2d"])
    assert np.abs(tif_oracle.warp_backward(tar.astype(np.float64), nw) - g["warp_f64"]).max() < FLOAT_TOL
    assert np.array_equal(tif_oracle.flow_warp_meshgrid(nw[:, :, 0] * 7.0, nw[:, :, 1] * 5.0), g["meshgrid"])
    assert np.array_equal(tif_oracle.erp_of_wraparound(g["wraparound_in"].copy()), g["wraparound_out"])


@pytest.mark.parametrize("name", SUMMARY)
def test_summary_cases(golden, name):
    g = golden(name)
    src, tar, flows = inputs(g)
    erp_h, wt, pad = int(g["erp_h"]), int(g["tangent_w"]), float(g["pad"])
    full = np.stack(tif_oracle.erp2ico_image(src, wt, pad, True)).astype(np.uint8)
    assert sha(full) == str(g["ico_full_sha"])
    assert np.array_equal(full[..., :3].reshape(20, -1).sum(1), g["ico_full_facesum"])
    assert bool(g["ico_full_alpha_all255"]) == bool((full[..., 3] == 255).all())
    large = erp_h > 1100          # keep the CPU suite to a few minutes: one blend for the 4K cases
    for blend in (("straightforward",) if large else ("straightforward", "normwarp")):
        tag = blend[:2]
        f = tif_oracle.ico2erp_flow(flows, erp_h, pad, src, tar, True, blend)
        stride = int(g["flow_stride"])
        assert np.abs(f.reshape(-1, 2)[::stride] - g["flow_%s_samples" % tag]).max() < FLOAT_TOL
        assert np.allclose(f.reshape(-1, 2).sum(0), g["flow_%s_sum" % tag], rtol=1e-9, atol=1e-6)
        assert np.abs(np.stack((f[0], f[erp_h - 1])) - g["flow_%s_poles" % tag]).max() < FLOAT_TOL
        if blend == "normwarp":
            w = tif_oracle.warp_backward(tar, f)
            # the flow differs from the reference's by ~1e-13, which can move a uint8 rounding: compare samples
            diff = (w.reshape(-1)[::int(g["warp_stride"])].astype(int) - g["warp_samples"].astype(int))
            assert np.abs(diff).max() <= 1 and (diff != 0).mean() < 1e-3
    faces = oracle_membership(flows, erp_h, pad)
    assert np.array_equal([d["candidates"] for d in faces], g["candidates"])
    assert np.array_equal([d["accepted"] for d in faces], g["accepted"])
    assert np.array_equal([d["bbox"] for d in faces], g["bbox"])
    for i, d in enumerate(faces):
        packed = np.stack((d["rows"], d["cols"], d["ix"], d["iy"], d["mult"])).astype(np.int32)
        assert sha(packed) == str(g["mem_sha"][i]), "face %d membership differs" % i


def test_survey_known_answers():
    """SURVEY.md Appendix C (H=1024, seed 0): values measured by the survey on the live reference."""
    src, tar, flows = synth.synth(1024, 0, 480)
    assert sha(src)[:12] == "36cc2019a264" and sha(tar)[:12] == "aa8949c7a1d5" and sha(np.stack(flows))[:12] == "963319776cbf"
    faces = oracle_membership(flows, 1024, 0.3)
    assert [d["accepted"] for d in faces][:6] == [386412, 385705, 385754, 385705, 386412, 173840]
    assert [d["candidates"] for d in faces][:6] == [902000] * 5 + [310351]
    assert faces[5]["bbox"] == (-104, 512, 254, 756) and faces[0]["bbox"] == (-1, 2048, -1, 438)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_against_live_reference():
    ref = ref_shim.load_reference()
    pi, fw = ref["projection_icosahedron"], ref["flow_warp"]
    erp_h, wt, pad = 160, 72, 0.3
    src, tar = synth.synth_images(erp_h, 33)
    flows = synth.synth_flows(wt, 20, 9.0)
    for full in (True, False):
        a = pi.erp2ico_image(src, wt, pad, full)
        b = tif_oracle.erp2ico_image(src, wt, pad, full)
        assert all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(a, b))
    for blend, tol in (("straightforward", 0.0), ("normwarp", 1e-11)):
        for wrap in (False, True):
            fa = pi.ico2erp_flow(flows, erp_h, pad, src, tar, wrap, blend)
            fb = tif_oracle.ico2erp_flow(flows, erp_h, pad, src, tar, wrap, blend)
            assert np.abs(fa - fb).max() <= tol
    assert np.array_equal(fw.warp_backward(tar, fa), tif_oracle.warp_backward(tar, fa))
    assert np.array_equal(fw.flow_warp_meshgrid(fa[..., 0] * 9, fa[..., 1] * 9), tif_oracle.flow_warp_meshgrid(fa[..., 0] * 9, fa[..., 1] * 9))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_on_demo_pair_config1():
    """BASELINE.json config #1: the shipped 1280x640 demo pair, pad 0.3, Wt 480 (resample stage; the
    stitch of this size is covered by summary_h640)."""
    from PIL import Image
    import os
    d = os.path.join(os.path.dirname(ref_shim.REFERENCE_SRC), "data", "replica_360", "hotel_0")
    img = np.asarray(Image.open(os.path.join(d, "0000_rgb_pano.jpg")))
    assert img.shape == (640, 1280, 3)
    ref = ref_shim.load_reference()
    a = ref["projection_icosahedron"].erp2ico_image(img, 480, 0.3, True)
    b = tif_oracle.erp2ico_image(img, 480, 0.3, True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_erp2ico_float_images_against_live_reference():
    """Float ERP images (projection_icosahedron.py:231-241: sampled in float64, cast into the int64 tangent arrays)."""
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(21)
    for dtype in (np.float64, np.float32):
        img = rng.uniform(0.0, 255.0, (64, 128, 3)).astype(dtype)
        for full in (True, False):
            a = ref["projection_icosahedron"].erp2ico_image(img, 32, 0.3, full)
            b = tif_oracle.erp2ico_image(img, 32, 0.3, full)
            assert all(x.dtype == y.dtype and np.array_equal(x, y) for x, y in zip(a, b)), (dtype, full)
    with pytest.raises(IndexError):
        ref["projection_icosahedron"].erp2ico_image(rng.uniform(0, 1, (64, 128)), 32, 0.3)
