"""Sweep HostPipeline chunk / slot settings (pinned host in/out, 32 pairs of 2048x1024)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as G
G.build(); pkg = G.load_package()
dev = torch.device("cuda", 0)
B, H, WT = 32, 1024, 480
plan = pkg._geometry.get_plan(0.3, H, WT, dev)
g = torch.Generator(device=dev); g.manual_seed(0)
src = torch.randint(0, 256, (B, H, 2 * H, 3), dtype=torch.uint8, device=dev, generator=g).cpu().pin_memory()
tar = torch.randint(0, 256, (B, H, 2 * H, 3), dtype=torch.uint8, device=dev, generator=g).cpu().pin_memory()
flows = (torch.rand((B, plan.tangent_pixels, 2), device=dev, generator=g) * 4 - 2).cpu().pin_memory()
out = torch.empty((B, H, 2 * H, 2), dtype=torch.float32).pin_memory()
ts = torch.empty((B, plan.tangent_pixels, 4), dtype=torch.uint8).pin_memory()
tt = torch.empty((B, plan.tangent_pixels, 4), dtype=torch.uint8).pin_memory()
for chunk, slots in ((4, 3), (8, 3), (4, 4), (2, 4), (8, 2), (16, 2)):
    pipe = pkg.pipeline.HostPipeline(0.3, H, WT, "normwarp", True, chunk=chunk, slots=slots, device=dev)
    for _ in range(2):
        pipe.run(src, tar, flows, out, ts, tt)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        pipe.run(src, tar, flows, out, ts, tt)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    gb = B * (pipe.h2d_bytes_per_pair() + pipe.d2h_bytes_per_pair()) / 1e9
    print("chunk %2d slots %d: %.1f ms  %.0f pairs/s  (%.1f GB/s both ways)" % (chunk, slots, dt * 1e3, B / dt, gb / dt))
    del pipe
