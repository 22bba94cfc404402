"""Drop-in replacement of the hot functions of the reference's src/projection_cubemap.py.

    get_cubemap_parameters(padding_size=0.0)
    erp2cubemap_image(erp_image_mat, padding_size=0.0, face_image_size=None)
    cubemap2erp_flow(cubemap_flows_list, face_flows_wraparound=None, erp_image_height=None, padding_size=0.0,
                     image_erp_src=None, image_erp_tar=None, wrap_around=False)

Same kernels as the icosahedron stage (k_erp2ico, k_lift, k_gather) over a 6-face plan of kind TIF_PLAN_CUBEMAP:
square faces, the cubemap stage's own resampling grid and ERP bounding boxes, weights 0.95 / ||warp error||_2 on
uint8-rounded samples (or the inner-square indicator without images), no seam fix.  numpy in -> numpy out with the
reference's types: six float64 [S,S,3] face images, float64 [H,2H,2] flow.
"""
import numpy as np
import torch

from . import _geometry, _lib, _ops
from ._geometry import CUBEMAP_TANGENT_POINTS, cubemap_face_erp_range


def _device():
    if not torch.cuda.is_available():
        raise _lib.TifError("no CUDA device: the cubemap stage has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def get_cubemap_parameters(padding_size=0.0):
    """Tangent points, the 4 spherical corners of every face and the padded face ranges
    (src/projection_cubemap.py:21-124)."""
    lat = np.arctan(np.sqrt(2.0) * 0.5)
    q = np.pi
    corners = np.zeros((6, 4, 2), dtype=float)
    corners[0] = [[0.25 * q, lat], [0.75 * q, lat], [0.75 * q, -lat], [0.25 * q, -lat]]
    corners[1] = [[-0.75 * q, lat], [-0.25 * q, lat], [-0.25 * q, -lat], [-0.75 * q, -lat]]
    corners[2] = [[-0.75 * q, lat], [-0.75 * q, lat], [0.25 * q, lat], [-0.25 * q, lat]]
    corners[3] = [[-0.25 * q, -lat], [0.25 * q, -lat], [0.75 * q, -lat], [-0.75 * q, -lat]]
    corners[4] = [[-0.25 * q, lat], [0.25 * q, lat], [0.25 * q, -lat], [-0.25 * q, -lat]]
    corners[5] = [[0.75 * q, lat], [-0.75 * q, lat], [-0.75 * q, -lat], [0.75 * q, -lat]]
    return {"tangent_points": np.array(CUBEMAP_TANGENT_POINTS, dtype=float), "face_points": corners,
            "face_erp_range": cubemap_face_erp_range(padding_size)}


def face_meshgrid(face_erp_range_sphere_list, erp_image_height):
    """ERP pixel grid of the bounding box of one face, wrapped into the image (src/projection_cubemap.py:190-243).
    Host geometry (the device plan evaluates the same box, `_geometry.cubemap_erp_bounding_box`)."""
    x_min, x_max, y_min, y_max = _geometry.cubemap_erp_bounding_box(face_erp_range_sphere_list, erp_image_height)
    face_erp_x, face_erp_y = np.meshgrid(np.linspace(x_min, x_max, x_max - x_min + 1), np.linspace(y_min, y_max, y_max - y_min + 1))
    return np.remainder(face_erp_x, erp_image_height * 2), np.remainder(face_erp_y, erp_image_height)


def erp2cubemap_image(erp_image_mat, padding_size=0.0, face_image_size=None):
    """ERP image -> 6 padded cube faces (+x, -x, +y, -y, +z, -z), src/projection_cubemap.py:127-187."""
    erp_image_mat = np.asarray(erp_image_mat)
    if erp_image_mat.ndim != 3 or erp_image_mat.shape[2] != 3 or erp_image_mat.dtype != np.uint8:
        raise NotImplementedError("erp2cubemap_image: only uint8 [H,W,3] images are on the accelerated path")
    erp_h, erp_w = erp_image_mat.shape[:2]
    if face_image_size is None:
        face_image_size = int(erp_w / 4.0)
    dev = _device()
    plan = _geometry.get_plan(padding_size, erp_h, face_image_size, dev, 0, _lib.PLAN_CUBEMAP)
    erp_dev = torch.from_numpy(np.ascontiguousarray(erp_image_mat)).to(dev)[None]
    out = _ops.ops.erp2ico(erp_dev, _ops.register_plan(plan), True)
    stack = out.view(6, face_image_size, face_image_size, 4)[..., :3].cpu().numpy()
    return [stack[f].astype(np.float64) for f in range(6)]        # the reference fills float arrays with the uint8 samples


def cubemap2erp_flow(cubemap_flows_list, face_flows_wraparound=None, erp_image_height=None, padding_size=0.0,
                     image_erp_src=None, image_erp_tar=None, wrap_around=False):
    """Stitch the 6 cube-face flows into an ERP flow, src/projection_cubemap.py:246-357.  Like the reference,
    `wrap_around` is accepted and never used (no seam fix in this stage)."""
    if face_flows_wraparound is not None:
        raise NotImplementedError("cubemap2erp_flow: face_flows_wraparound is not on the accelerated path")
    if len(cubemap_flows_list) != 6:
        raise ValueError("the cubemap images number is not 6")
    size = np.shape(cubemap_flows_list[0])[0]
    erp_h = int(size * 2.0) if erp_image_height is None else int(erp_image_height)
    dev = _device()
    plan = _geometry.get_plan(padding_size, erp_h, size, dev, 0, _lib.PLAN_CUBEMAP)
    for f in cubemap_flows_list:
        if tuple(np.shape(f)) != (size, size, 2):
            raise ValueError("cube face flow of shape %s, expected %s" % (np.shape(f), (size, size, 2)))
    flows = np.stack([np.asarray(f, dtype=np.float32) for f in cubemap_flows_list]).reshape(1, plan.tangent_pixels, 2)
    flows_dev = torch.from_numpy(flows).to(dev)
    blend, src_dev, tar_dev = _lib.BLEND_CUBEMAP_INSIDE, None, None
    if image_erp_src is not None and image_erp_tar is not None:
        imgs = []
        for img in (image_erp_src, image_erp_tar):
            img = np.asarray(img)
            if img.dtype != np.uint8 or img.shape != (plan.erp_h, plan.erp_w, 3):
                raise NotImplementedError("cubemap2erp_flow needs uint8 [H,W,3] ERP images of the flow's size")
            imgs.append(torch.from_numpy(np.ascontiguousarray(img)).to(dev)[None])
        src_dev, tar_dev = imgs
        blend = _lib.BLEND_CUBEMAP_WARP_ERROR
    out = _ops.ops.ico2erp_flow(flows_dev, src_dev, tar_dev, _ops.register_plan(plan), blend, False)
    return out[0].cpu().numpy().astype(np.float64)
