#!/bin/bash
# usage (GPU box): scripts/ab_warp.sh [variant ...]   -- warp_backward kernel time at 2048x1024 and 3840x1920 with the regular
# library and with each prebuilt variant (see scripts/ab_variants.sh)
run() {
python - "$1" <<'PY'
import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as e
e.build(); pkg = e.load_package(); lib = pkg._lib.load()
dev = torch.device("cuda", 0)
for h in (1024, 1920):
    B = 16
    g = torch.Generator(device=dev); g.manual_seed(1)
    img = torch.randint(0, 256, (B, h, 2 * h, 3), dtype=torch.uint8, device=dev, generator=g)
    yy = torch.arange(h, device=dev).view(1, h, 1).float(); xx = torch.arange(2 * h, device=dev).view(1, 1, 2 * h).float()
    flow = torch.stack((4 * torch.sin(xx / 97.0 + yy / 41.0) + 1.5, 3 * torch.cos(yy / 53.0) - 0.5 + 0 * xx), -1).expand(B, h, 2 * h, 2).contiguous()
    flow = flow + torch.rand((B, 1, 1, 2), device=dev, generator=g)
    out = torch.empty_like(img)
    s = torch.cuda.current_stream().cuda_stream
    def run():
        lib.tif_warp_backward_u8(img.data_ptr(), flow.data_ptr(), 0, B, h, 2 * h, 3, 0, out.data_ptr(), s)
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("[%s] h=%d  %.3f ms per %d images  frac %.3f" % (sys.argv[1], h, ms, B, B * h * 2 * h * 14 / (ms * 1e-3) / 6543.7e9), flush=True)
PY
}
unset TIF_B200_LIB
run base
for n in "$@"; do
  export TIF_B200_LIB=$PWD/360opticalflow-tangentimages_b200/build/variants/$n.so
  run $n
done
