"""HostPipeline (pinned host buffers in, pinned host buffers out) against the device-resident calls: every tangent
format gives the same bytes, with the device slots reused across chunks and calls."""
import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("blend", ["normwarp", "straightforward"])
def test_host_pipeline_matches_device_calls(pkg, blend):
    erp_h, wt, pad, batch = 128, 64, 0.3, 5
    dev = torch.device("cuda", 0)
    plan = pkg._geometry.get_plan(pad, erp_h, wt, dev)
    rng = np.random.default_rng(3)
    src = torch.from_numpy(rng.integers(0, 256, (batch, erp_h, 2 * erp_h, 3), dtype=np.uint8)).pin_memory()
    tar = torch.from_numpy(rng.integers(0, 256, (batch, erp_h, 2 * erp_h, 3), dtype=np.uint8)).pin_memory()
    one = np.concatenate([f.reshape(-1, 2) for f in synth.synth_flows(wt)]).astype(np.float32)
    flows = torch.from_numpy(one[None] + rng.uniform(-0.5, 0.5, (batch, 1, 2)).astype(np.float32)).contiguous().pin_memory()

    ops, lib = pkg._ops, pkg._lib
    code = lib.BLEND_NORMWARP if blend == "normwarp" else lib.BLEND_STRAIGHTFORWARD
    d_src, d_tar, d_flows = src.to(dev), tar.to(dev), flows.to(dev)
    tan_s = torch.empty((batch, plan.tangent_pixels, 4), dtype=torch.uint8, device=dev)
    tan_t = torch.empty_like(tan_s)
    want = torch.empty((batch, erp_h, 2 * erp_h, 2), dtype=torch.float32, device=dev)
    scratch = torch.empty(max(int(plan.lib.tif_ico2erp_scratch_bytes(plan.handle, batch, code)), 8), dtype=torch.uint8, device=dev)
    ops.erp2ico_out(plan, d_src, True, tan_s)
    ops.erp2ico_out(plan, d_tar, True, tan_t)
    ops.ico2erp_flow_out(plan, d_flows, d_src, d_tar, code, True, want, scratch)
    torch.cuda.synchronize()

    for fmt in ("rgb", "gray", "rgba"):
        pipe = pkg.pipeline.HostPipeline(pad, erp_h, wt, blend, True, chunk=2, slots=2, device=dev, tangent_format=fmt)
        out = torch.empty(want.shape, dtype=torch.float32).pin_memory()
        ts = torch.empty(pipe.tangent_out_shape(batch), dtype=torch.uint8).pin_memory()
        tt = torch.empty(pipe.tangent_out_shape(batch), dtype=torch.uint8).pin_memory()
        for _ in range(2):          # the slots are reused: nothing of an earlier chunk may leak
            pipe.run(src, tar, flows, out, ts, tt)
        assert torch.equal(out, want.cpu()), fmt
        if fmt == "rgba":
            assert torch.equal(ts, tan_s.cpu()) and torch.equal(tt, tan_t.cpu())
        elif fmt == "rgb":
            assert torch.equal(ts, tan_s.cpu()[..., :3]) and torch.equal(pipe.expand_rgba(tt), tan_t.cpu())
        else:
            c = tan_s.cpu().to(torch.int64)         # cv2 COLOR_BGR2GRAY on the stack's first three channels
            gray = (3735 * c[..., 0] + 19235 * c[..., 1] + 9798 * c[..., 2] + (1 << 14)) >> 15
            assert torch.equal(ts.to(torch.int64), gray)
        assert pipe.h2d_bytes_per_pair() == 2 * erp_h * 2 * erp_h * 3 + plan.tangent_pixels * 8
