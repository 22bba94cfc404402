"""Batched host-to-host pipeline: the call a user of the reference makes, at production batch sizes.

The reference API takes and returns host (numpy) arrays, one frame pair per call.  `HostPipeline`
keeps that contract -- host buffers in, host buffers out -- for a batch of pairs: pinned host memory,
host->device copies, kernels and device->host copies run on three CUDA streams over a ring of device
slots, so PCIe traffic in both directions overlaps the kernels chunk by chunk.

Per pair the device receives src + tar (2 x H*W*3 bytes) and the tangent flows (P*8 bytes, produced
by the per-face estimator on the host) and returns the stitched ERP flow (H*W*8 bytes) and both
tangent-image stacks, the estimator's input.  The path is PCIe-bound (the kernels are ~10x faster than
the link), so the stacks travel in the most compact form the consumer accepts:

  tangent_format="rgb"   (default) [B,P,3]: the alpha channel is a constant 255 for full face images
                         (projection_icosahedron.py:244) and is not sent; `expand_rgba` restores it on the host
  tangent_format="gray"  [B,P]: cv2.cvtColor(img[:, :, :3], COLOR_BGR2GRAY), the first thing the reference's DIS
                         wrapper computes (flow_estimate.py:31-47), evaluated on the device bit for bit
  tangent_format="rgba"  [B,P,4]: the stack as the kernels write it
"""
import torch

from . import _geometry, _lib, _ops

_FORMATS = {"rgba": 4, "rgb": 3, "gray": 1}


class HostPipeline:
    def __init__(self, padding_size, erp_height, tangent_width, face_blending_method="normwarp", wrap_around=True,
                 chunk=4, slots=3, device=None, return_tangent_images=True, tangent_format="rgb"):
        if face_blending_method not in ("straightforward", "normwarp"):
            raise ValueError(face_blending_method)
        if tangent_format not in _FORMATS:
            raise ValueError("tangent_format must be one of %s" % sorted(_FORMATS))
        self.plan = _geometry.get_plan(padding_size, erp_height, tangent_width, device)
        self.device = self.plan.device
        self.blend = _lib.BLEND_NORMWARP if face_blending_method == "normwarp" else _lib.BLEND_STRAIGHTFORWARD
        self.blend_name = face_blending_method
        self.wrap_around = bool(wrap_around)
        self.chunk, self.n_slots = int(chunk), int(slots)
        self.return_tangent = bool(return_tangent_images)
        self.tangent_format = tangent_format
        self.tangent_channels = _FORMATS[tangent_format]
        p = self.plan
        dev = self.device
        if tangent_format != "rgba" and p.tangent_pixels % 4 != 0:
            raise ValueError("packed tangent formats need a tangent stack of a multiple of 4 pixels per image")
        with torch.cuda.device(dev):
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
            self.slots = []
            for _ in range(self.n_slots):
                nbytes = max(int(p.lib.tif_ico2erp_scratch_bytes(p.handle, self.chunk, self.blend)), 8)
                slot = {
                    "src": torch.empty((self.chunk, p.erp_h, p.erp_w, 3), dtype=torch.uint8, device=dev),
                    "tar": torch.empty((self.chunk, p.erp_h, p.erp_w, 3), dtype=torch.uint8, device=dev),
                    "flows": torch.empty((self.chunk, p.tangent_pixels, 2), dtype=torch.float32, device=dev),
                    "tan_src": torch.empty((self.chunk, p.tangent_pixels, 4), dtype=torch.uint8, device=dev),
                    "tan_tar": torch.empty((self.chunk, p.tangent_pixels, 4), dtype=torch.uint8, device=dev),
                    "out": torch.empty((self.chunk, p.erp_h, p.erp_w, 2), dtype=torch.float32, device=dev),
                    "scratch": torch.empty(nbytes, dtype=torch.uint8, device=dev),
                    "ev_in": torch.cuda.Event(), "ev_run": torch.cuda.Event(), "ev_free": torch.cuda.Event(),
                }
                if tangent_format != "rgba" and self.return_tangent:
                    shape = self.tangent_out_shape(self.chunk)
                    slot["pack_src"] = torch.empty(shape, dtype=torch.uint8, device=dev)
                    slot["pack_tar"] = torch.empty(shape, dtype=torch.uint8, device=dev)
                self.slots.append(slot)
        self.kernel_launches = 0

    def tangent_out_shape(self, batch):
        p = self.plan
        return (batch, p.tangent_pixels) if self.tangent_channels == 1 else (batch, p.tangent_pixels, self.tangent_channels)

    def describe(self):
        return ("pipeline.HostPipeline.run (pinned host in/out; D2H returns the ERP flow and both tangent stacks as %s)"
                % self.tangent_format)

    def h2d_bytes_per_pair(self):
        p = self.plan
        return 2 * p.erp_h * p.erp_w * 3 + p.tangent_pixels * 8

    def d2h_bytes_per_pair(self):
        p = self.plan
        return p.erp_h * p.erp_w * 8 + (2 * p.tangent_pixels * self.tangent_channels if self.return_tangent else 0)

    @staticmethod
    def expand_rgba(rgb):
        """[..., 3] uint8 (host or device) -> [..., 4] with the constant alpha 255 of a full face image."""
        out = rgb.new_full(rgb.shape[:-1] + (4,), 255)
        out[..., :3] = rgb
        return out

    def _pack(self, s, n):
        p = self.plan
        fmt = _lib.PACK_RGB if self.tangent_format == "rgb" else _lib.PACK_GRAY_CV
        for a, b in (("tan_src", "pack_src"), ("tan_tar", "pack_tar")):
            _lib.check(p.lib.tif_pack_tangent_u8(s[a].data_ptr(), n * p.tangent_pixels, fmt, s[b].data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream), "tif_pack_tangent_u8")
        return 2

    def run(self, src, tar, flows, out_flow, out_tan_src=None, out_tan_tar=None):
        """All arguments are (preferably pinned) HOST tensors: src/tar uint8 [B,H,W,3], flows float32 [B,P,2],
        out_flow float32 [B,H,W,2], out_tan_* uint8 of `tangent_out_shape(B)`.  Returns when every output is on the host."""
        p = self.plan
        batch = src.shape[0]
        want_tan = self.return_tangent and out_tan_src is not None
        if want_tan and (tuple(out_tan_src.shape) != self.tangent_out_shape(batch) or tuple(out_tan_tar.shape) != self.tangent_out_shape(batch)):
            raise ValueError("tangent outputs must be uint8 %s for tangent_format=%r" % (self.tangent_out_shape(batch), self.tangent_format))
        launches = 0
        with torch.cuda.device(self.device):
            for i, b0 in enumerate(range(0, batch, self.chunk)):
                n = min(self.chunk, batch - b0)
                s = self.slots[i % self.n_slots]
                with torch.cuda.stream(self.s_in):
                    if i >= self.n_slots:
                        self.s_in.wait_event(s["ev_free"])
                    s["src"][:n].copy_(src[b0:b0 + n], non_blocking=True)
                    s["tar"][:n].copy_(tar[b0:b0 + n], non_blocking=True)
                    s["flows"][:n].copy_(flows[b0:b0 + n], non_blocking=True)
                    s["ev_in"].record(self.s_in)
                with torch.cuda.stream(self.s_run):
                    self.s_run.wait_event(s["ev_in"])
                    _ops.erp2ico_out(p, s["src"][:n], True, s["tan_src"][:n])
                    _ops.erp2ico_out(p, s["tar"][:n], True, s["tan_tar"][:n])
                    packed = want_tan and self.tangent_format != "rgba"
                    if packed:
                        launches += self._pack(s, n)
                    _ops.ico2erp_flow_out(p, s["flows"][:n], s["src"][:n], s["tar"][:n], self.blend, self.wrap_around,
                                          s["out"][:n], s["scratch"])
                    launches += 4 if self.blend == _lib.BLEND_STRAIGHTFORWARD else 5
                    s["ev_run"].record(self.s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(s["ev_run"])
                    out_flow[b0:b0 + n].copy_(s["out"][:n], non_blocking=True)
                    if want_tan:
                        if self.tangent_format == "rgba":
                            out_tan_src[b0:b0 + n].copy_(s["tan_src"][:n], non_blocking=True)
                            out_tan_tar[b0:b0 + n].copy_(s["tan_tar"][:n], non_blocking=True)
                        else:
                            out_tan_src[b0:b0 + n].copy_(s["pack_src"][:n], non_blocking=True)
                            out_tan_tar[b0:b0 + n].copy_(s["pack_tar"][:n], non_blocking=True)
                    s["ev_free"].record(self.s_out)
            self.s_out.synchronize()
        self.kernel_launches += launches
        return out_flow
