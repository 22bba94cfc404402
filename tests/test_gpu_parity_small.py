"""GPU parity on the committed small fixtures (outputs of the live reference) and against the oracle."""
import numpy as np
import pytest

from oracle import synth, tif_oracle

pytestmark = pytest.mark.gpu

SMALL = ["small_h64_noise", "small_h80_smooth_seam", "small_h48_pad01"]
FLOW_TOL = 1e-3     # px, north_star tolerance for ERP flow (float32 path vs float64 reference)


def flow_mismatch(got, want, src, tar, flows, g, wrap, blend):
    """Pixels whose flow differs from the reference by more than FLOW_TOL, not counting pixels where the
    reference's own map is discontinuous within tolerance of a contributing end point (from the oracle):
      - an end point within 1e-5 rad of the +-pi seam may land on either side (u jumps by W).  With
        wrap_around the seam fix normally hides that, but the reference applies its modulo twice
        (sph_coord_modulo, then erp_sph_modulo inside sph2erp) and an end point within an ulp of +pi is
        +pi for the seam test and -pi for the pixel coordinate, so its own result is off by W there;
      - normwarp:   scipy mode='constant' reads 255 outright outside [0, n-1], so an end point within
                    1e-3 px of that edge switches one face's weight."""
    err = np.abs(got - want).max(axis=2)
    _, dbg = tif_oracle.ico2erp_flow(flows, int(g["erp_h"]), float(g["pad"]), src, tar, wrap, blend, return_debug=True)
    excuse = np.zeros(err.shape, dtype=bool)
    excuse |= dbg["seam_dist"] < 1e-5
    if not wrap:
        # without the seam fix the faces of a pixel can disagree by a whole image width (one end point wrapped,
        # another not); the reference's weighted mean of such values is meaningless and magnifies any 1e-7
        # difference in the weights by W
        excuse |= dbg["fu_spread"] > 0.25 * want.shape[1]
    if blend == "normwarp":
        excuse |= dbg["edge_dist"] < 1e-3
    return (err > FLOW_TOL) & ~excuse


def _inputs(g):
    src, tar = synth.synth_images(int(g["erp_h"]), int(g["seed"]), bool(g["smooth"]))
    flows = synth.synth_flows(int(g["tangent_w"]), 20, float(g["u_offset"]))
    return src, tar, flows


@pytest.mark.parametrize("name", SMALL)
def test_erp2ico_matches_reference_fixture(pkg, golden, name):
    g = golden(name)
    src, tar, flows = _inputs(g)
    pico = pkg.projection_icosahedron
    full = pico.erp2ico_image(src, int(g["tangent_w"]), float(g["pad"]), True)
    assert len(full) == 20 and all(t.dtype == np.int64 and t.shape[2] == 4 for t in full)
    assert np.array_equal(np.stack(full), g["ico_full"].astype(np.int64))
    part = pico.erp2ico_image(src, int(g["tangent_w"]), float(g["pad"]), False)
    assert np.array_equal(np.stack(part), g["ico_part"].astype(np.int64))


@pytest.mark.parametrize("name", SMALL)
def test_membership_bit_exact(pkg, golden, name):
    g = golden(name)
    plan = pkg._geometry.get_plan(float(g["pad"]), int(g["erp_h"]), int(g["tangent_w"]))
    count, entries = plan.export_stage2()
    count = count.cpu().numpy()
    entries = entries.cpu().numpy().view(np.uint32)
    h, w = count.shape
    rows, cols = np.nonzero(np.ones((h, w), bool))
    got = []
    for k in range(entries.shape[0]):
        sel = count.reshape(-1) > k
        e = entries[k].reshape(-1)[sel]
        got.append(np.stack((e >> 25, rows[sel], cols[sel], e & 2047, (e >> 11) & 2047, (e >> 22) & 7), 1))
    got = np.concatenate(got).astype(np.int64)
    got = got[np.lexsort((got[:, 2], got[:, 1], got[:, 0]))]
    want = np.stack((g["mem_face"], g["mem_row"], g["mem_col"], g["mem_ix"], g["mem_iy"], g["mem_mult"]), 1).astype(np.int64)
    want = want[np.lexsort((want[:, 2], want[:, 1], want[:, 0]))]
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("blend", ["straightforward", "normwarp"])
@pytest.mark.parametrize("wrap", [False, True])
def test_ico2erp_flow_matches_reference_fixture(pkg, golden, name, blend, wrap):
    g = golden(name)
    src, tar, flows = _inputs(g)
    got = pkg.projection_icosahedron.ico2erp_flow(flows, int(g["erp_h"]), float(g["pad"]), src, tar, wrap, blend)
    want = g["flow_%s_%d" % (blend[:2], int(wrap))]
    assert got.dtype == np.float64 and got.shape == want.shape
    bad = flow_mismatch(got, want, src, tar, flows, g, wrap, blend)
    assert bad.sum() == 0, "%d pixels beyond %g px" % (bad.sum(), FLOW_TOL)


@pytest.mark.parametrize("name", SMALL)
def test_warp_meshgrid_wraparound(pkg, golden, name):
    g = golden(name)
    src, tar, flows = _inputs(g)
    nw = g["flow_no_1"]
    fw, fp = pkg.flow_warp, pkg.flow_postproc
    assert np.array_equal(fw.warp_backward(tar, nw), g["warp_u8"])
    out2d = fw.warp_backward(tar[:, :, 0].copy(), nw)
    assert out2d.shape == g["warp_2d"].shape and np.array_equal(out2d, g["warp_2d"])
    wf = fw.warp_backward(tar.astype(np.float64), nw)
    assert wf.dtype == np.float64 and np.array_equal(wf, g["warp_f64"])      # float64 images stay float64: exact
    mg = fw.flow_warp_meshgrid(nw[:, :, 0] * 7.0, nw[:, :, 1] * 5.0)
    assert mg.shape == g["meshgrid"].shape and np.array_equal(mg, g["meshgrid"])
    big = g["wraparound_in"].copy()
    out = fp.erp_of_wraparound(big)
    assert np.array_equal(out, g["wraparound_out"])
    assert np.array_equal(big[:, :, 0], g["wraparound_out"][:, :, 0])      # argument's u plane is rewritten


@pytest.mark.parametrize("smooth,offset", [(False, 0.0), (True, 25.0)])
def test_oracle_agrees_at_mid_size(pkg, smooth, offset):
    """GPU vs oracle at H=256: IID-noise images (the harshest case for the photometric weights), and a
    band-limited pair with a 25 px flow offset that pushes end points across the seam and the image edge."""
    erp_h, wt, pad = 256, 120, 0.3
    src, tar = synth.synth_images(erp_h, 21, smooth)
    flows = synth.synth_flows(wt, 20, offset)
    pico = pkg.projection_icosahedron
    got = pico.erp2ico_image(tar, wt, pad, True)
    want = tif_oracle.erp2ico_image(tar, wt, pad, True)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    for blend in ("straightforward", "normwarp"):
        g = pico.ico2erp_flow(flows, erp_h, pad, src, tar, True, blend)
        w, dbg = tif_oracle.ico2erp_flow(flows, erp_h, pad, src, tar, True, blend, return_debug=True)
        excuse = dbg["seam_dist"] < 1e-5
        if blend == "normwarp":
            excuse |= dbg["edge_dist"] < 1e-3
        bad = ((np.abs(g - w).max(axis=2) > FLOW_TOL) & ~excuse).sum()
        assert bad == 0, (blend, bad, np.abs(g - w).max())


@pytest.mark.parametrize("erp_h,wt", [(90, 50), (77 * 2, 37), (66, 95), (88, 53), (45, 31)])
def test_sizes_that_are_not_tile_multiples(pkg, erp_h, wt):
    """Edge sizes: tangent width not a multiple of the 8-wide warp tile, ERP height not a multiple of 4, a tangent
    image wider than the ERP is tall.  ERP rows of 6 H bytes are 16-byte multiples only for H = 88: that case takes
    the shared-memory staged resampler with partial 16 x 8 patches, the others the direct gathers (H = 45: an odd
    height, images that are not multiples of 4 bytes).  uint8 exact,
    membership implied by the flow parity."""
    src, tar = synth.synth_images(erp_h, 13)
    flows = synth.synth_flows(wt, 20, 2.0)
    pico = pkg.projection_icosahedron
    for full in (True, False):
        got = pico.erp2ico_image(src, wt, 0.3, full)
        want = tif_oracle.erp2ico_image(src, wt, 0.3, full)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
    for blend in ("straightforward", "normwarp"):
        want, dbg = tif_oracle.ico2erp_flow(flows, erp_h, 0.3, src, tar, True, blend, return_debug=True)
        got = pico.ico2erp_flow(flows, erp_h, 0.3, src, tar, True, blend)
        excuse = dbg["seam_dist"] < 1e-5
        if blend == "normwarp":
            excuse |= dbg["edge_dist"] < 1e-3
        assert (((np.abs(got - want).max(axis=2)) > FLOW_TOL) & ~excuse).sum() == 0
    assert np.array_equal(pkg.flow_warp.warp_backward(tar, want), tif_oracle.warp_backward(tar, want))


@pytest.mark.parametrize("pad", [0.0, 0.05, 0.6])
def test_padding_extremes(pkg, pad):
    """No padding (faces only touch, pixels on shared edges rely on the polygon epsilon) and a padding large enough
    for most pixels to be seen by five or more faces."""
    erp_h, wt = 64, 40
    src, tar = synth.synth_images(erp_h, 21, smooth=True)
    flows = synth.synth_flows(wt, 20, 1.0)
    pico = pkg.projection_icosahedron
    for full in (True, False):
        got = pico.erp2ico_image(src, wt, pad, full)
        want = tif_oracle.erp2ico_image(src, wt, pad, full)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
    for blend in ("straightforward", "normwarp"):
        want, dbg = tif_oracle.ico2erp_flow(flows, erp_h, pad, src, tar, True, blend, return_debug=True)
        got = pico.ico2erp_flow(flows, erp_h, pad, src, tar, True, blend)
        excuse = dbg["seam_dist"] < 1e-5
        if blend == "normwarp":
            excuse |= dbg["edge_dist"] < 1e-3
        assert (((np.abs(got - want).max(axis=2)) > FLOW_TOL) & ~excuse).sum() == 0, (pad, blend)


@pytest.mark.parametrize("scale", [6.0, 20.0])
def test_large_flows(pkg, scale):
    """Tangent flows of tens of pixels on 60-pixel tangent images: end points leave their face, cross the seam and
    the poles, and the lift leaves its small-angle fast path.  Smooth images, so that a 1e-4 px sampling difference
    does not turn into a different weight."""
    erp_h, wt = 128, 60
    src, tar = synth.synth_images(erp_h, 5, smooth=True)
    flows = [f * np.float32(scale) for f in synth.synth_flows(wt, 20, 0.5)]
    pico = pkg.projection_icosahedron
    for wrap in (True, False):
        for blend in ("straightforward", "normwarp"):
            want, dbg = tif_oracle.ico2erp_flow(flows, erp_h, 0.3, src, tar, wrap, blend, return_debug=True)
            got = pico.ico2erp_flow(flows, erp_h, 0.3, src, tar, wrap, blend)
            excuse = dbg["seam_dist"] < 1e-5
            if not wrap:
                excuse |= dbg["fu_spread"] > 0.25 * want.shape[1]
            if blend == "normwarp":
                excuse |= dbg["edge_dist"] < 1e-3
            bad = ((np.abs(got - want).max(axis=2)) > FLOW_TOL) & ~excuse
            assert bad.sum() == 0, (scale, wrap, blend, int(bad.sum()), float(np.abs(got - want)[bad].max()))


def test_erp2ico_image_api_corners(pkg):
    """Inputs the estimator never feeds but the reference accepts (projection_icosahedron.py:178-245): float images are
    sampled in float64 and truncated into the int64 tangent arrays; a single-channel map fails in the reference with an
    IndexError (it writes channel 3 of a 1-channel array), so it does here; RGBA input loses its alpha."""
    pico = pkg.projection_icosahedron
    rng = np.random.default_rng(21)
    erp_h, wt = 64, 32
    for dtype in (np.float64, np.float32):
        img = rng.uniform(0.0, 255.0, (erp_h, 2 * erp_h, 3)).astype(dtype)
        for full in (True, False):
            got = pico.erp2ico_image(img, wt, 0.3, full)
            want = tif_oracle.erp2ico_image(img, wt, 0.3, full)
            assert len(got) == 20 and all(g.dtype == np.int64 and np.array_equal(g, w) for g, w in zip(got, want)), (dtype, full)
    with pytest.raises(IndexError):
        pico.erp2ico_image(rng.uniform(0, 1, (erp_h, 2 * erp_h)), wt, 0.3)
    u8 = rng.integers(0, 256, (erp_h, 2 * erp_h, 4), dtype=np.uint8)
    assert all(np.array_equal(a, b) for a, b in zip(pico.erp2ico_image(u8, wt, 0.3), pico.erp2ico_image(u8[:, :, :3].copy(), wt, 0.3)))
