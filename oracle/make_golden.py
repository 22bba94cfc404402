"""Generate tests/golden/*.npz from the LIVE reference (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference is mounted:

    python oracle/make_golden.py

The reference has no golden vectors of its own, so these fixtures -- outputs of the
unmodified reference imported through oracle/ref_shim.py on inputs from oracle/synth.py -- are the
pin for oracle/tif_oracle.py and, through it, for the CUDA path.  Float fixtures are compared with
a 1e-9 tolerance (numpy's arcsin/arctan2 differ in the last ulp between CPU dispatch targets);
integer / uint8 fixtures and hashes are compared exactly.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
PADS = (0.0, 0.1, 0.3, 0.4)
BBOX_HEIGHTS = (64, 128, 640, 1024, 1920, 2048)
SAMPLES = 20000  # strided samples kept per large array


def stride_for(n):
    """Odd stride that keeps about SAMPLES entries of an n-entry array (>= 61)."""
    return max(61, (n // SAMPLES) | 1)


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_bbox(sc, area, erp_h):
    """Reference integer bbox (projection_icosahedron.py:299-305) through the reference's own sph2erp."""
    cs, rs = sc.sph2erp(area[0], area[2], erp_h, sph_modulo=False)
    ce, re = sc.sph2erp(area[1], area[3], erp_h, sph_modulo=False)
    cs = int(cs) if int(cs) > 0 else int(cs - 0.5)
    ce = int(ce + 0.5) if int(ce) > 0 else int(ce)
    rs = int(rs) if int(rs) > 0 else int(rs - 0.5)
    re = int(re + 0.5) if int(re) > 0 else int(re)
    return cs, ce, rs, re


def ref_membership(ref, face, erp_h, pad, tangent_w, tangent_h):
    """Accepted (row, col, ix, iy) of one face by replaying the reference's own calls
    (projection_icosahedron.py:285-329); duplicates from wrapped columns are kept, in reference order."""
    pi, sc, gp, poly, fp = (ref[k] for k in ("projection_icosahedron", "spherical_coordinates",
                                              "gnomonic_projection", "polygon", "flow_postproc"))
    par = pi.get_icosahedron_parameters(face, pad)
    th0, ph0 = par["tangent_point"]
    tri = np.array(par["triangle_points_tangent"])
    xmin, xmax = np.amin(tri[:, 0]), np.amax(tri[:, 0])
    ymin, ymax = np.amin(tri[:, 1]), np.amax(tri[:, 1])
    rng_gnom = [xmin, xmax, ymin, ymax]
    eps = abs(xmax - xmin) / (2 * tangent_w)
    cs, ce, rs, re = ref_bbox(sc, par["availied_ERP_area"], erp_h)
    xv, yv = np.meshgrid(np.linspace(cs, ce, ce - cs + 1), np.linspace(rs, re, re - rs + 1))
    xv, yv = fp.erp_pixles_modulo(xv, yv, 2 * erp_h, erp_h)
    sph = sc.erp2sph((xv, yv), erp_h, sph_modulo=False)
    gx, gy, _ = gp.gnomonic_projection(sph[0], sph[1], th0, ph0, True)
    acc = poly.inside_polygon_2d(np.stack((gx.flatten(), gy.flatten()), axis=1), tri, on_line=True,
                                 eps=eps).reshape(gx.shape)
    ix, iy = gp.gnomonic2pixel(gx[acc], gy[acc], 0.0, tangent_w, tangent_h, rng_gnom)
    return (yv[acc].astype(np.int64), xv[acc].astype(np.int64), ix.astype(np.int64), iy.astype(np.int64),
            int(acc.size), (cs, ce, rs, re))


def unique_membership(rows, cols, ix, iy, erp_w):
    """Collapse duplicate wrapped pixels; return arrays sorted by (row, col) + multiplicity."""
    key = rows * erp_w + cols
    uk, first, cnt = np.unique(key, return_index=True, return_counts=True)
    return rows[first], cols[first], ix[first], iy[first], cnt


def face_tables(ref):
    pi, sc = ref["projection_icosahedron"], ref["spherical_coordinates"]
    out = {"pads": np.array(PADS), "bbox_heights": np.array(BBOX_HEIGHTS)}
    tp = np.zeros((len(PADS), 20, 2))
    tri = np.zeros((len(PADS), 20, 3, 2))
    area = np.zeros((len(PADS), 20, 4))
    bbox = np.zeros((len(PADS), len(BBOX_HEIGHTS), 20, 4), dtype=np.int64)
    for pidx, pad in enumerate(PADS):
        for f in range(20):
            par = pi.get_icosahedron_parameters(f, pad)
            tp[pidx, f] = par["tangent_point"]
            tri[pidx, f] = np.array(par["triangle_points_tangent"])
            area[pidx, f] = par["availied_ERP_area"]
            for hidx, h in enumerate(BBOX_HEIGHTS):
                bbox[pidx, hidx, f] = ref_bbox(sc, par["availied_ERP_area"], h)
    out.update(tangent_point=tp, tri=tri, area=area, bbox=bbox)
    np.savez_compressed(os.path.join(OUT, "face_table.npz"), **out)


def small_case(ref, name, erp_h, tangent_w, pad, seed, smooth, u_offset):
    """Full arrays: everything a parity test needs for a configuration that runs in well under a second."""
    pi, fw, fp = ref["projection_icosahedron"], ref["flow_warp"], ref["flow_postproc"]
    tangent_h = synth.tangent_height(tangent_w)
    src, tar = synth.synth_images(erp_h, seed, smooth)
    flows = synth.synth_flows(tangent_w, 20, u_offset)
    out = dict(erp_h=erp_h, tangent_w=tangent_w, tangent_h=tangent_h, pad=pad, seed=seed,
               smooth=int(smooth), u_offset=u_offset, sha_src=sha(src), sha_tar=sha(tar), sha_flows=sha(np.stack(flows)))
    full = pi.erp2ico_image(src, tangent_w, pad, True)
    assert all(t.dtype == np.int64 and t.shape == (tangent_h, tangent_w, 4) for t in full)
    out["ico_full"] = np.stack(full).astype(np.uint8)
    part = pi.erp2ico_image(src, tangent_w, pad, False)
    out["ico_part"] = np.stack(part).astype(np.uint8)
    for blend in ("straightforward", "normwarp"):
        for wrap in (False, True):
            f = pi.ico2erp_flow(flows, erp_h, pad, src, tar, wrap, blend)
            assert f.dtype == np.float64
            out["flow_%s_%d" % (blend[:2], int(wrap))] = f
    rows_l, cols_l, ix_l, iy_l, m_l, face_l, cand, bbox = [], [], [], [], [], [], [], []
    for face in range(20):
        r, c, ix, iy, ncand, bb = ref_membership(ref, face, erp_h, pad, tangent_w, tangent_h)
        r, c, ix, iy, m = unique_membership(r, c, ix, iy, 2 * erp_h)
        rows_l.append(r); cols_l.append(c); ix_l.append(ix); iy_l.append(iy); m_l.append(m)
        face_l.append(np.full(r.shape, face)); cand.append(ncand); bbox.append(bb)
    out.update(mem_face=np.concatenate(face_l).astype(np.int16), mem_row=np.concatenate(rows_l).astype(np.int16),
               mem_col=np.concatenate(cols_l).astype(np.int16), mem_ix=np.concatenate(ix_l).astype(np.int16),
               mem_iy=np.concatenate(iy_l).astype(np.int16), mem_mult=np.concatenate(m_l).astype(np.int8),
               candidates=np.array(cand), bbox=np.array(bbox))
    nw = out["flow_no_1"]
    out["warp_u8"] = fw.warp_backward(tar, nw)
    out["warp_f64"] = fw.warp_backward(tar.astype(np.float64), nw)
    out["warp_2d"] = fw.warp_backward(tar[:, :, 0].copy(), nw)
    out["meshgrid"] = fw.flow_warp_meshgrid(nw[:, :, 0] * 7.0, nw[:, :, 1] * 5.0)
    big = np.stack((nw[:, :, 0] * 40.0, nw[:, :, 1]), -1)
    out["wraparound_in"] = big.copy()
    out["wraparound_out"] = fp.erp_of_wraparound(big.copy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "written")


def summary_case(ref, name, erp_h, tangent_w, pad, seed, smooth, u_offset, warp=True, images=None):
    """Hashes, counts, sums and strided float samples for the configurations too large to commit in full.
    `images` = (src, tar) replaces the synthetic pair (BASELINE.json config #1: the shipped demo pair); the fixture
    then carries the decoded images themselves, because the GPU box has no /root/reference to read them from."""
    pi, fw = ref["projection_icosahedron"], ref["flow_warp"]
    tangent_h = synth.tangent_height(tangent_w)
    erp_w = 2 * erp_h
    src, tar = synth.synth_images(erp_h, seed, smooth) if images is None else images
    flows = synth.synth_flows(tangent_w, 20, u_offset)
    out = dict(erp_h=erp_h, tangent_w=tangent_w, tangent_h=tangent_h, pad=pad, seed=seed, smooth=int(smooth),
               u_offset=u_offset, sha_src=sha(src), sha_tar=sha(tar), sha_flows=sha(np.stack(flows)))
    if images is not None:
        out.update(image_src=src, image_tar=tar)
        out["ico_tar_sha"] = sha(np.stack(pi.erp2ico_image(tar, tangent_w, pad, True)).astype(np.uint8))
    full = np.stack(pi.erp2ico_image(src, tangent_w, pad, True)).astype(np.uint8)
    out["ico_full_sha"] = sha(full)
    out["ico_full_facesum"] = full[..., :3].reshape(20, -1).sum(1)
    out["ico_full_alpha_all255"] = int((full[..., 3] == 255).all())
    out["ico_stride"] = stride_for(full.size)
    out["ico_full_samples"] = full.reshape(-1)[::stride_for(full.size)]
    for blend in ("straightforward", "normwarp"):
        f = pi.ico2erp_flow(flows, erp_h, pad, src, tar, True, blend)
        tag = blend[:2]
        out["flow_stride"] = stride_for(erp_h * erp_w)
        out["flow_%s_samples" % tag] = f.reshape(-1, 2)[::stride_for(erp_h * erp_w)].copy()
        out["flow_%s_sum" % tag] = f.reshape(-1, 2).sum(0)
        out["flow_%s_absmean" % tag] = np.abs(f).reshape(-1, 2).mean(0)
        out["flow_%s_poles" % tag] = np.stack((f[0], f[erp_h - 1]))          # the two quirk rows, in full
        # where the reference's own map is discontinuous (an end point on the +-pi seam, or on the hard edge of the
        # mode='constant' sampling under normwarp) a float32 implementation may land on the other side: the oracle
        # identifies those pixels, the fixture keeps the flags of the strided samples
        from oracle import tif_oracle
        fo, dbg = tif_oracle.ico2erp_flow(flows, erp_h, pad, src, tar, True, blend, return_debug=True)
        assert np.abs(fo - f).max() < 1e-9, (name, blend, np.abs(fo - f).max())
        excuse = dbg["seam_dist"] < 1e-5
        if blend == "normwarp":
            excuse |= dbg["edge_dist"] < 1e-3
        out["flow_%s_excuse" % tag] = excuse.reshape(-1)[::stride_for(erp_h * erp_w)].copy()
        if warp and blend == "normwarp":
            w = fw.warp_backward(tar, f)
            out["warp_sha"] = sha(w)
            out["warp_sum"] = int(w.sum())
            out["warp_stride"] = stride_for(w.size)
            out["warp_samples"] = w.reshape(-1)[::stride_for(w.size)]
    cand, acc, bbox, hashes = [], [], [], []
    for face in range(20):
        r, c, ix, iy, ncand, bb = ref_membership(ref, face, erp_h, pad, tangent_w, tangent_h)
        acc.append(r.size)
        r, c, ix, iy, m = unique_membership(r, c, ix, iy, erp_w)
        cand.append(ncand); bbox.append(bb)
        hashes.append(sha(np.stack((r, c, ix, iy, m)).astype(np.int32)))
    out.update(candidates=np.array(cand), accepted=np.array(acc), bbox=np.array(bbox), mem_sha=np.array(hashes))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "written")


def demo_pair():
    """The reference's shipped demo pair (data/replica_360/hotel_0, 1280x640), decoded as main.py does
    (image_io.image_read = PIL)."""
    from PIL import Image
    d = os.path.join(os.path.dirname(ref_shim.REFERENCE_SRC), "data", "replica_360", "hotel_0")
    pair = [np.asarray(Image.open(os.path.join(d, n))) for n in ("0000_rgb_pano.jpg", "0001_rgb_pano.jpg")]
    assert all(p.shape == (640, 1280, 3) and p.dtype == np.uint8 for p in pair)
    return pair[0], pair[1]


def cubemap_flows(size, seed):
    """Smooth per-face flows for the cubemap stage, float32 [S,S,2] x 6."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    flows = []
    for f in range(6):
        a, b = rng.uniform(1, 3, 2)
        u = 3.0 * np.sin(2 * np.pi * (xx / size + f / 6.0)) + a
        v = 2.0 * np.cos(2 * np.pi * (yy / size - f / 6.0)) - b
        flows.append(np.stack((u, v), -1).astype(np.float32))
    return flows


def cubemap_case(name, erp_h, pad, seed, smooth, full):
    """Outputs of the live reference's cubemap stage (projection_cubemap.py): full arrays for small sizes,
    hashes + strided samples otherwise."""
    import importlib
    ref_shim.load_reference()
    pc = importlib.import_module("projection_cubemap")
    src, tar = synth.synth_images(erp_h, seed, smooth)
    faces = pc.erp2cubemap_image(src, pad)
    size = faces[0].shape[0]
    assert all(f.dtype == np.float64 and f.shape == (size, size, 3) for f in faces)
    flows = cubemap_flows(size, seed)
    out = dict(erp_h=erp_h, pad=pad, seed=seed, smooth=int(smooth), size=size, sha_src=sha(src), sha_tar=sha(tar),
               sha_flows=sha(np.stack(flows)))
    stack = np.stack(faces)
    assert np.array_equal(stack, np.round(stack)) and stack.min() >= 0 and stack.max() <= 255
    plain = pc.cubemap2erp_flow(flows, None, erp_h, pad, None, None, True)
    weighted = pc.cubemap2erp_flow(flows, None, erp_h, pad, src, tar, True)
    if full:
        out.update(faces=stack.astype(np.uint8), flow_plain=plain, flow_weighted=weighted)
    else:
        st = stride_for(erp_h * 2 * erp_h)
        out.update(faces_sha=sha(stack.astype(np.uint8)), flow_stride=st, flow_plain_samples=plain.reshape(-1, 2)[::st].copy(),
                   flow_weighted_samples=weighted.reshape(-1, 2)[::st].copy(), flow_plain_sum=plain.reshape(-1, 2).sum(0))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "written")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load_reference()
    if "--demo-only" in sys.argv:
        summary_case(ref, "summary_demo_pair_h640", 640, 480, 0.3, -1, False, 0.0, images=demo_pair())
        return
    if "--cubemap-only" in sys.argv:
        cubemap_case("cubemap_h64_pad04", 64, 0.4, 3, False, True)
        cubemap_case("cubemap_h96_pad00_smooth", 96, 0.0, 5, True, True)
        cubemap_case("cubemap_summary_h1024", 1024, 0.4, 0, False, False)
        return
    face_tables(ref)
    small_case(ref, "small_h64_noise", 64, 32, 0.3, 3, False, 0.0)
    small_case(ref, "small_h80_smooth_seam", 80, 40, 0.3, 5, True, 12.0)
    small_case(ref, "small_h48_pad01", 48, 24, 0.1, 7, False, 0.0)
    summary_case(ref, "summary_h1024_noise", 1024, 480, 0.3, 0, False, 0.0)
    summary_case(ref, "summary_h640_smooth_seam", 640, 480, 0.3, 11, True, 40.0)
    summary_case(ref, "summary_demo_pair_h640", 640, 480, 0.3, -1, False, 0.0, images=demo_pair())
    cubemap_case("cubemap_h64_pad04", 64, 0.4, 3, False, True)
    cubemap_case("cubemap_h96_pad00_smooth", 96, 0.0, 5, True, True)
    cubemap_case("cubemap_summary_h1024", 1024, 0.4, 0, False, False)
    if "--large" in sys.argv:
        summary_case(ref, "summary_h1920_noise", 1920, 480, 0.3, 1, False, 0.0)
        summary_case(ref, "summary_h2048_smooth", 2048, 480, 0.3, 2, True, 0.0)


if __name__ == "__main__":
    main()
