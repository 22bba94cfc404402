"""The call sequence of PanoOpticalFlow.estimate (src/flow_estimate.py:111-196) -- ERP step, cubemap step, icosahedron
step, each followed by the global-rotation warp of the target image, then the rotation of the flow end points back --
with the per-image flow estimator (OpenCV DIS in the reference, a fixed stage outside this package) replaced by the
same analytic flow on both sides.  The oracle runs the sequence on the CPU, the package on the GPU: every step of
`estimate` except DIS composes, and the final flow agrees."""
import numpy as np
import pytest

from oracle import synth, tif_oracle

pytestmark = pytest.mark.gpu


def fake_flow(shape, phase):
    """Stands in for `optical_flow_base_line_method(src, tar)`: a smooth float32 flow of the image's shape."""
    h, w = shape[0], shape[1]
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    u = 2.5 * np.sin(2 * np.pi * (xx / w + phase)) + 1.0
    v = 1.5 * np.cos(2 * np.pi * (yy / h - phase)) - 0.25
    return np.stack((u, v), -1).astype(np.float32)


def estimate(mods, src, tar, tangent_w, pad_ico, pad_cube):
    """flow_estimate.py:122-196 with the three steps enabled (the class default enables only the last)."""
    proj_ico, proj_cm, flow_warp, projection = mods
    erp_h = src.shape[0]
    rotations = []
    # 0) ERP step
    flow_erp = fake_flow(src.shape, 0.1).astype(np.float64)
    tar, rot = flow_warp.global_rotation_warping(tar, flow_erp, False, "3D")
    rotations.append(rot)
    # 1) cubemap step
    faces_src, faces_tar = proj_cm.erp2cubemap_image(src, pad_cube), proj_cm.erp2cubemap_image(tar, pad_cube)
    flows = [fake_flow(faces_src[i].shape, 0.2 + i / 6.0) for i in range(6)]
    flow_cube = proj_cm.cubemap2erp_flow(flows, erp_h, pad_cube, src, tar)
    tar, rot = flow_warp.global_rotation_warping(tar, flow_cube, False, "3D")
    rotations.append(rot)
    # 2) icosahedron step
    faces_src, faces_tar = proj_ico.erp2ico_image(src, tangent_w, pad_ico, True), proj_ico.erp2ico_image(tar, tangent_w, pad_ico, True)
    flows = [fake_flow(faces_src[i].shape, 0.3 + i / 20.0) for i in range(20)]
    flow = proj_ico.ico2erp_flow(flows, erp_h, pad_ico, src, tar, True, "normwarp")
    # 3) back through the rotations
    for rot in reversed(rotations):
        flow = projection.flow_rotate_endpoint(flow, rot.T)
    return flow, tar, rotations


class OracleCubemap:
    erp2cubemap_image = staticmethod(tif_oracle.erp2cubemap_image)

    @staticmethod
    def cubemap2erp_flow(flows, erp_h, pad, src, tar):
        return tif_oracle.cubemap2erp_flow(flows, erp_h, pad, src, tar)


class OracleIco:
    erp2ico_image = staticmethod(tif_oracle.erp2ico_image)

    @staticmethod
    def ico2erp_flow(flows, erp_h, pad, src, tar, wrap, blend):
        return tif_oracle.ico2erp_flow(flows, erp_h, pad, src, tar, wrap, blend)


class OracleWarp:
    @staticmethod
    def global_rotation_warping(img, flow, forward, kind):
        return tif_oracle.global_rotation_warping(img, flow, forward, kind)


class OracleProjection:
    flow_rotate_endpoint = staticmethod(tif_oracle.flow_rotate_endpoint)


class PkgCubemap:
    def __init__(self, pkg):
        self.m = pkg.projection_cubemap

    def erp2cubemap_image(self, img, pad):
        return self.m.erp2cubemap_image(img, pad)

    def cubemap2erp_flow(self, flows, erp_h, pad, src, tar):
        return self.m.cubemap2erp_flow(flows, erp_image_height=erp_h, padding_size=pad, image_erp_src=src, image_erp_tar=tar, wrap_around=True)


def test_estimate_sequence(pkg):
    erp_h, wt = 96, 40
    src, tar = synth.synth_images(erp_h, 17, smooth=True)
    want, tar_w, rots_w = estimate((OracleIco, OracleCubemap, OracleWarp, OracleProjection), src, tar.copy(), wt, 0.3, 0.1)
    got, tar_g, rots_g = estimate((pkg.projection_icosahedron, PkgCubemap(pkg), pkg.flow_warp, pkg.projection), src, tar.copy(), wt, 0.3, 0.1)
    for a, b in zip(rots_g, rots_w):
        assert np.abs(a - b).max() < 1e-9
    assert tar_g.dtype == np.uint8 and (tar_g != tar_w).mean() < 1e-3           # twice-rotated target image
    err = np.abs(got - want)
    err[..., 0] = np.minimum(err[..., 0], np.abs(err[..., 0] - 2 * erp_h))       # the +-W/2 convention at the seam
    err = err.max(axis=2)
    # the few target pixels that rounded differently after two rotations move a normwarp weight by ~1e-3
    assert np.median(err) < 1e-5 and (err > 2e-3).mean() < 2e-3, (float(np.median(err)), float((err > 2e-3).mean()), float(err.max()))
