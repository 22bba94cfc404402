"""BASELINE.json config #5: level-1 (80-face) subdivision -- an extension the reference does not have
(SURVEY.md Appendix E).  Parity is self-consistency: GPU == the oracle's composition of the reference's
primitives (bit-exact membership / indices / uint8 samples, flow <= 1e-3 px), every ERP pixel covered,
and the same machinery at level 0 == the reference (tests/test_gpu_parity_*.py)."""
import numpy as np
import pytest
import torch

from oracle import synth, tif_oracle

pytestmark = pytest.mark.gpu
PAD = 0.3


def level1_flows(wt, heights, seed=0):
    rng = np.random.default_rng(seed)
    flows = []
    for f, ht in enumerate(heights):
        yy, xx = np.meshgrid(np.arange(ht), np.arange(wt), indexing="ij")
        u = 3.0 * np.sin(2 * np.pi * (xx / wt + f / 80.0)) + rng.uniform(-1, 1)
        v = 2.0 * np.cos(2 * np.pi * (yy / ht - f / 80.0)) + rng.uniform(-1, 1)
        flows.append(np.stack((u, v), -1).astype(np.float32))
    return flows


@pytest.mark.parametrize("erp_h,wt", [(256, 64), (1024, 480)])
def test_level1_against_oracle(pkg, erp_h, wt):
    """GPU == oracle for the 80-face subdivision, at a small size and at 2048x1024 with the default 480-pixel faces."""
    src, tar = synth.synth_images(erp_h, 4)
    plan = pkg._geometry.get_plan(PAD, erp_h, wt, level=1)
    assert plan.n_faces == 80 and len(set(plan.face_heights)) == 4
    heights = [tif_oracle.face_tangent_height(tif_oracle.subdivided_face(f, PAD), wt, 1) for f in range(80)]
    assert heights == plan.face_heights
    pico = pkg.projection_icosahedron
    got = pico.erp2ico_image(src, wt, PAD, True, subdivision_level=1)
    want = tif_oracle.erp2ico_image(src, wt, PAD, True, level=1)
    assert len(got) == 80 and all(np.array_equal(a, b) for a, b in zip(got, want))
    part = pico.erp2ico_image(src, wt, PAD, False, subdivision_level=1)
    want_part = tif_oracle.erp2ico_image(src, wt, PAD, False, level=1)
    assert all(np.array_equal(a, b) for a, b in zip(part, want_part))
    flows = level1_flows(wt, heights)
    # membership + indices, bit exact
    _, dbg = tif_oracle.ico2erp_flow(flows, erp_h, PAD, None, None, True, "straightforward", return_debug=True, level=1)
    count, entries = plan.export_stage2()
    count = count.cpu().numpy().reshape(-1)
    entries = entries.cpu().numpy().view(np.uint32).reshape(entries.shape[0], -1)
    got_rows = []
    for k in range(entries.shape[0]):
        pix = np.nonzero(count > k)[0]
        e = entries[k][pix]
        got_rows.append(np.stack((e >> 25, pix // plan.erp_w, pix % plan.erp_w, e & 2047, (e >> 11) & 2047, (e >> 22) & 7), 1))
    got_rows = np.concatenate(got_rows).astype(np.int64)
    got_rows = got_rows[np.lexsort((got_rows[:, 2], got_rows[:, 1], got_rows[:, 0]))]
    want_rows = np.concatenate([np.stack((np.full(d["rows"].shape, f), d["rows"], d["cols"], d["ix"], d["iy"], d["mult"]), 1)
                                for f, d in enumerate(dbg["faces"])]).astype(np.int64)
    want_rows = want_rows[np.lexsort((want_rows[:, 2], want_rows[:, 1], want_rows[:, 0]))]
    assert np.array_equal(got_rows, want_rows)
    assert count.min() >= 1                      # every ERP pixel is covered by at least one face
    for blend in ("straightforward", "normwarp"):
        want_f, dbg = tif_oracle.ico2erp_flow(flows, erp_h, PAD, src, tar, True, blend, return_debug=True, level=1)
        got_f = pico.ico2erp_flow(flows, erp_h, PAD, src, tar, True, blend, subdivision_level=1)
        excuse = dbg["seam_dist"] < 1e-5
        if blend == "normwarp":
            excuse |= dbg["edge_dist"] < 1e-3
        err = np.abs(got_f - want_f).max(axis=2)
        assert ((err > 1e-3) & ~excuse).sum() == 0, (blend, err.max())


def test_config5_4096x2048_80_faces(pkg):
    """Full size of config #5: the plan matches SURVEY.md section 8d's work counts, every pixel is covered, the
    stitch is deterministic and finite, and a constant image resamples to itself."""
    erp_h, wt = 2048, 480
    plan = pkg._geometry.get_plan(PAD, erp_h, wt, level=1)
    assert plan.n_faces == 80 and sorted(set(plan.face_heights)) == [346, 416, 485, 502]
    assert plan.tangent_pixels == 34980 * 480
    assert plan.accepted_unique > 3.1 * erp_h * 2 * erp_h           # SURVEY: 26.9 M accepted incl. duplicates (3.209 x HW)
    count, _ = plan.export_stage2()
    assert int(count.min()) >= 1
    dev = plan.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    src = torch.randint(0, 256, (1, erp_h, 2 * erp_h, 3), dtype=torch.uint8, device=dev, generator=gen)
    tar = torch.randint(0, 256, (1, erp_h, 2 * erp_h, 3), dtype=torch.uint8, device=dev, generator=gen)
    flows = (torch.rand((1, plan.tangent_pixels, 2), device=dev, generator=gen) * 6.0 - 3.0).contiguous()
    ops, lib = pkg._ops, pkg._lib
    out = torch.empty((1, erp_h, 2 * erp_h, 2), dtype=torch.float32, device=dev)
    out2 = torch.empty_like(out)
    for blend in (lib.BLEND_STRAIGHTFORWARD, lib.BLEND_NORMWARP):
        ops.ico2erp_flow_out(plan, flows, src, tar, blend, True, out)
        ops.ico2erp_flow_out(plan, flows, src, tar, blend, True, out2)
        assert torch.equal(out, out2) and bool(torch.isfinite(out).all())
        # away from the poles a |flow| <= 3 tangent px stays a small ERP flow
        mid = out[0, erp_h // 4: 3 * erp_h // 4]
        assert float(mid.abs().max()) < 40.0
    const = torch.full((1, erp_h, 2 * erp_h, 3), 201, dtype=torch.uint8, device=dev)
    tan = pkg.projection_icosahedron.erp2ico_image(const, wt, PAD, subdivision_level=1)
    assert tan.shape == (1, plan.tangent_pixels, 4) and bool((tan[..., :3] == 201).all())
    pkg._geometry.clear_plans()
