"""Batched host-to-host pipeline: the call a user of the reference makes, at production batch sizes.

The reference API takes and returns host (numpy) arrays, one frame pair per call.  `HostPipeline`
keeps that contract -- host buffers in, host buffers out -- for a batch of pairs: pinned host memory,
host->device copies, kernels and device->host copies run on three CUDA streams over a ring of device
slots, so PCIe traffic in both directions overlaps the kernels chunk by chunk.

Per pair the device receives src + tar (2 x H*W*3 bytes) and the tangent flows (P*8 bytes, produced
by the per-face estimator on the host) and returns both tangent-image stacks (2 x P*4 bytes, the
estimator's input) and the stitched ERP flow (H*W*8 bytes).
"""
import torch

from . import _geometry, _lib, _ops


class HostPipeline:
    def __init__(self, padding_size, erp_height, tangent_width, face_blending_method="normwarp", wrap_around=True,
                 chunk=4, slots=3, device=None, return_tangent_images=True):
        if face_blending_method not in ("straightforward", "normwarp"):
            raise ValueError(face_blending_method)
        self.plan = _geometry.get_plan(padding_size, erp_height, tangent_width, device)
        self.device = self.plan.device
        self.blend = _lib.BLEND_NORMWARP if face_blending_method == "normwarp" else _lib.BLEND_STRAIGHTFORWARD
        self.wrap_around = bool(wrap_around)
        self.chunk, self.n_slots = int(chunk), int(slots)
        self.return_tangent = bool(return_tangent_images)
        p = self.plan
        dev = self.device
        with torch.cuda.device(dev):
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
            self.slots = []
            for _ in range(self.n_slots):
                nbytes = max(int(p.lib.tif_ico2erp_scratch_bytes(p.handle, self.chunk, _lib.BLEND_NORMWARP)), 8)
                self.slots.append({
                    "src": torch.empty((self.chunk, p.erp_h, p.erp_w, 3), dtype=torch.uint8, device=dev),
                    "tar": torch.empty((self.chunk, p.erp_h, p.erp_w, 3), dtype=torch.uint8, device=dev),
                    "flows": torch.empty((self.chunk, p.tangent_pixels, 2), dtype=torch.float32, device=dev),
                    "tan_src": torch.empty((self.chunk, p.tangent_pixels, 4), dtype=torch.uint8, device=dev),
                    "tan_tar": torch.empty((self.chunk, p.tangent_pixels, 4), dtype=torch.uint8, device=dev),
                    "out": torch.empty((self.chunk, p.erp_h, p.erp_w, 2), dtype=torch.float32, device=dev),
                    "scratch": torch.empty(nbytes, dtype=torch.uint8, device=dev),
                    "ev_in": torch.cuda.Event(), "ev_run": torch.cuda.Event(), "ev_free": torch.cuda.Event(),
                })
        self.kernel_launches = 0

    def h2d_bytes_per_pair(self):
        p = self.plan
        return 2 * p.erp_h * p.erp_w * 3 + p.tangent_pixels * 8

    def d2h_bytes_per_pair(self):
        p = self.plan
        return p.erp_h * p.erp_w * 8 + (2 * p.tangent_pixels * 4 if self.return_tangent else 0)

    def run(self, src, tar, flows, out_flow, out_tan_src=None, out_tan_tar=None):
        """All arguments are (preferably pinned) HOST tensors: src/tar uint8 [B,H,W,3], flows float32 [B,P,2],
        out_flow float32 [B,H,W,2], out_tan_* uint8 [B,P,4].  Returns when every output is on the host."""
        p = self.plan
        batch = src.shape[0]
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
                    _ops.ico2erp_flow_out(p, s["flows"][:n], s["src"][:n], s["tar"][:n], self.blend, self.wrap_around,
                                          s["out"][:n], s["scratch"])
                    launches += 4 if self.blend == _lib.BLEND_STRAIGHTFORWARD else 6
                    s["ev_run"].record(self.s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(s["ev_run"])
                    out_flow[b0:b0 + n].copy_(s["out"][:n], non_blocking=True)
                    if self.return_tangent and out_tan_src is not None:
                        out_tan_src[b0:b0 + n].copy_(s["tan_src"][:n], non_blocking=True)
                        out_tan_tar[b0:b0 + n].copy_(s["tan_tar"][:n], non_blocking=True)
                    s["ev_free"].record(self.s_out)
            self.s_out.synchronize()
        self.kernel_launches += launches
        return out_flow
