"""Drop-in replacement of the reference's src/projection_icosahedron.py hot functions.

Same names, argument meaning and return types as the reference:

    get_icosahedron_parameters(triangle_index, padding_size=0.0)
    erp2ico_image(erp_image, tangent_image_width, padding_size=0.0, full_face_image=True)
    ico2erp_flow(tangent_flows_list, erp_flow_height=None, padding_size=0.0, image_erp_src=None,
                 image_erp_tar=None, wrap_around=False, face_blending_method="straightforward")

numpy in -> numpy out exactly like the reference (list of 20 int64 [Ht,Wt,4] images; float64 [H,W,2]
flow).  Keyword-only extra: subdivision_level=1 selects the 80-face level-1 subdivision (an extension, the
reference stops at 20 faces; faces then have their own heights and the lists hold 80 arrays).  Torch CUDA tensors with a leading batch axis go straight through and stay on the device:
erp2ico_image([B,H,W,3] uint8) -> [B,20,Ht,Wt,4] uint8, ico2erp_flow([B,20,Ht,Wt,2] float32) ->
[B,H,W,2] float32.  All per-pixel work runs in libtif_b200.so; there is no CPU fallback.
"""
import logging

import numpy as np
import torch

from . import _geometry, _lib, _ops
from ._geometry import get_icosahedron_parameters, get_subdivided_parameters, tangent_image_height  # noqa: F401

log = logging.getLogger(__name__)

_BLENDS = {"straightforward": _lib.BLEND_STRAIGHTFORWARD, "normwarp": _lib.BLEND_NORMWARP}


def _device():
    if not torch.cuda.is_available():
        raise _lib.TifError("no CUDA device: the tangent-image hot path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _split_faces(stack, plan):
    """[..., P, C] flat tangent stack -> list of per-face [..., Ht_f, Wt, C] views (ragged at level 1)."""
    return [stack[..., o:o + h * plan.tangent_w, :].reshape(stack.shape[:-2] + (h, plan.tangent_w, stack.shape[-1]))
            for o, h in zip(plan.face_offsets, plan.face_heights)]


def erp2ico_image(erp_image, tangent_image_width, padding_size=0.0, full_face_image=True, *, subdivision_level=0):
    """Project an ERP image to the 20 padded tangent images (src/projection_icosahedron.py:163-248)."""
    if isinstance(erp_image, torch.Tensor):
        if erp_image.dim() != 4:
            raise ValueError("torch input must be [B,H,W,3] uint8")
        if erp_image.shape[-1] == 4:
            erp_image = erp_image[..., 0:3]
        plan = _geometry.get_plan(padding_size, erp_image.shape[1], tangent_image_width, erp_image.device, subdivision_level)
        out = _ops.ops.erp2ico(erp_image, _ops.register_plan(plan), bool(full_face_image))
        if subdivision_level:
            return out                    # ragged faces: flat [B,P,4]; plan.face_offsets / face_heights index it
        return out.view(erp_image.shape[0], plan.n_faces, plan.tangent_h, plan.tangent_w, 4)

    erp_image = np.asarray(erp_image)
    if erp_image.ndim == 3 and erp_image.shape[2] == 4:
        erp_image = erp_image[:, :, 0:3]
    if erp_image.ndim == 2:
        # the reference expands a single-channel map to [H,W,1], builds [Ht,Wt,1] tangent images and then writes their
        # channel 3 (projection_icosahedron.py:181-183,230-244): it never returns
        raise IndexError("index 3 is out of bounds for axis 2 with size 1")
    if erp_image.ndim != 3 or erp_image.shape[2] != 3:
        raise UnboundLocalError("tangent_image: the reference logs 'The channel number is %s' and fails" % (erp_image.shape[2:],))
    if erp_image.dtype != np.uint8:
        return _erp2ico_image_float(erp_image, tangent_image_width, padding_size, full_face_image, subdivision_level)
    if erp_image.shape[1] != erp_image.shape[0] * 2:
        log.error("the ERP image dimession is %s", np.shape(erp_image))     # the reference logs and continues
        raise ValueError("ERP image must be 2:1, got %s" % (erp_image.shape,))
    dev = _device()
    plan = _geometry.get_plan(padding_size, erp_image.shape[0], tangent_image_width, dev, subdivision_level)
    erp_dev = torch.from_numpy(np.ascontiguousarray(erp_image)).to(dev)[None]
    out = _ops.ops.erp2ico(erp_dev, _ops.register_plan(plan), bool(full_face_image))
    # the reference builds its images with np.full(..., 255): int64, 4 channels
    return [t.astype(np.int64) for t in _split_faces(out[0].cpu().numpy(), plan)]


def _erp2ico_image_float(erp_image, tangent_image_width, padding_size, full_face_image, subdivision_level):
    """Non-uint8 images: scipy samples them in float64 and the values are then cast into the reference's integer
    tangent arrays (np.full(..., 255) is int64; assignment truncates toward zero), projection_icosahedron.py:231-245."""
    if erp_image.shape[1] != erp_image.shape[0] * 2:
        log.error("the ERP image dimession is %s", np.shape(erp_image))
        raise ValueError("ERP image must be 2:1, got %s" % (erp_image.shape,))
    img = np.ascontiguousarray(erp_image, dtype=np.float32 if erp_image.dtype == np.float32 else np.float64)
    dev = _device()
    plan = _geometry.get_plan(padding_size, img.shape[0], tangent_image_width, dev, subdivision_level)
    img_dev = torch.from_numpy(img).to(dev)[None]
    out = torch.empty((1, plan.tangent_pixels, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(plan.lib.tif_erp2ico_float(plan.handle, img_dev.data_ptr(), _lib.DTYPE_F32 if img.dtype == np.float32 else _lib.DTYPE_F64,
                                              1, 3, int(bool(full_face_image)), out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "tif_erp2ico_float")
    vals = out[0].cpu().numpy()
    if img.dtype == np.float32:
        vals = vals.astype(np.float32)      # scipy's output takes the input's dtype before the integer cast
    rgb = np.trunc(vals).astype(np.int64)
    inside = plan.export_stage1()[2].cpu().numpy().astype(bool) if not full_face_image else np.ones(plan.tangent_pixels, dtype=bool)
    stack = np.concatenate((rgb, np.where(inside, 255, 0).astype(np.int64)[:, None]), axis=1)
    return [np.ascontiguousarray(t) for t in _split_faces(stack, plan)]


def ico2erp_flow(tangent_flows_list, erp_flow_height=None, padding_size=0.0, image_erp_src=None, image_erp_tar=None,
                 wrap_around=False, face_blending_method="straightforward", *, subdivision_level=0):
    """Stitch the 20 tangent flows into an ERP flow (src/projection_icosahedron.py:251-398)."""
    if face_blending_method not in _BLENDS:
        # the reference falls through to an UnboundLocalError on face_weight_mat
        raise UnboundLocalError("face_blending_method %r is not 'straightforward' or 'normwarp'" % (face_blending_method,))
    blend = _BLENDS[face_blending_method]

    if isinstance(tangent_flows_list, torch.Tensor):
        flows = tangent_flows_list
        if subdivision_level:
            # ragged faces: flat [B,P,2] stack; the tangent width cannot be read off the tensor
            raise NotImplementedError("level-1 torch input: call _ops.ico2erp_flow_out with a plan from "
                                      "_geometry.get_plan(..., level=1) and a flat [B,P,2] stack")
        if flows.dim() != 5 or flows.shape[-1] != 2:
            raise ValueError("torch input must be [B,20,Ht,Wt,2] float32")
        batch, n_faces, ht, wt = flows.shape[:4]
        erp_h = int(ht * 2.0) if erp_flow_height is None else int(erp_flow_height)
        plan = _geometry.get_plan(padding_size, erp_h, wt, flows.device)
        if n_faces != plan.n_faces or ht != plan.tangent_h:
            raise ValueError("expected [B,%d,%d,%d,2] tangent flows" % (plan.n_faces, plan.tangent_h, plan.tangent_w))
        return _ops.ops.ico2erp_flow(flows.reshape(batch, plan.tangent_pixels, 2), image_erp_src, image_erp_tar,
                                     _ops.register_plan(plan), blend, bool(wrap_around))

    n_expected = 20 * 4 ** int(subdivision_level)
    if len(tangent_flows_list) != n_expected:
        log.error("the ico face flow number is not %d", n_expected)
        raise ValueError("the ico face flow number is not %d" % n_expected)
    ht, wt = np.shape(tangent_flows_list[0])[:2]
    if np.shape(tangent_flows_list[0])[2] != 2:
        log.error("The flow channels number is %s", np.shape(tangent_flows_list[0])[2])
        raise ValueError("tangent flows must have 2 channels")
    erp_h = int(ht * 2.0) if erp_flow_height is None else int(erp_flow_height)
    dev = _device()
    plan = _geometry.get_plan(padding_size, erp_h, wt, dev, subdivision_level)
    for f, h in zip(tangent_flows_list, plan.face_heights):
        if tuple(np.shape(f)) != (h, wt, 2):
            raise ValueError("tangent flow of shape %s, expected %s" % (np.shape(f), (h, wt, 2)))
    flows = np.concatenate([np.asarray(f, dtype=np.float32).reshape(-1, 2) for f in tangent_flows_list])
    flows_dev = torch.from_numpy(flows).to(dev).reshape(1, plan.tangent_pixels, 2)
    src_dev = tar_dev = None
    if blend == _lib.BLEND_NORMWARP:
        imgs = []
        for img in (image_erp_src, image_erp_tar):
            img = np.asarray(img)
            if img.dtype != np.uint8 or img.shape != (plan.erp_h, plan.erp_w, 3):
                raise NotImplementedError("normwarp blending needs uint8 [H,W,3] ERP images of the flow's size "
                                          "(got %s %s)" % (img.dtype, img.shape))
            imgs.append(torch.from_numpy(np.ascontiguousarray(img)).to(dev)[None])
        src_dev, tar_dev = imgs
    out = _ops.ops.ico2erp_flow(flows_dev, src_dev, tar_dev, _ops.register_plan(plan), blend, bool(wrap_around))
    return out[0].cpu().numpy().astype(np.float64)
