"""Launch every hot kernel of the path a few times on one GPU (for `ncu`): python scripts/prof_kernels.py [batch] [erp_h]
Pass 1 warms up (plan build, first launches), pass 2 is the one to profile (`ncu -s <kernels in pass 1>`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import __graft_entry__ as entry
import bench

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
erp_h = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
entry.build()
pkg = entry.load_package()
ops, lib = pkg._ops, pkg._lib
dev = torch.device("cuda", 0)
plan = pkg._geometry.get_plan(bench.PAD, erp_h, bench.TANGENT_W, dev)
src, tar, flows = bench.synth_device(torch, plan, batch, 1234, dev)
tan = torch.empty((batch, plan.tangent_pixels, 4), dtype=torch.uint8, device=dev)
out = torch.empty((batch, plan.erp_h, plan.erp_w, 2), dtype=torch.float32, device=dev)
warped = torch.empty_like(tar)
scratch = torch.empty(int(plan.lib.tif_ico2erp_scratch_bytes(plan.handle, batch, lib.BLEND_NORMWARP)), dtype=torch.uint8, device=dev)
for _ in range(2):
    ops.erp2ico_out(plan, src, True, tan)                                                    # k_erp2ico_staged
    ops.ico2erp_flow_out(plan, flows, src, tar, lib.BLEND_NORMWARP, True, out, scratch)        # k_lift_nw, k_err_nw, k_nw_face_scale, k_blend_nw
    ops.ico2erp_flow_out(plan, flows, None, None, lib.BLEND_STRAIGHTFORWARD, True, out, scratch)   # k_lift<0>, k_gather<0>
    ops.ico2erp_flow_out(plan, flows, src, tar, lib.BLEND_NORMWARP, True, out, scratch)
    warped = ops.ops.warp_backward(tar, out, lib.WARP_WRAP)                                  # k_warp_rgb_fast
    torch.cuda.synchronize()
print("ok", float(out.abs().mean()), int(warped.sum()))
