"""Raw pinned-memory copy ceiling of this box: H2D alone, D2H alone, both at once (two streams), per rank.
Usage: python scripts/pcie_probe.py [MiB]   (torchrun for several ranks: every rank probes its own GPU concurrently)"""
import json
import os
import sys
import time

import torch


def probe(device, mib=256, reps=8):
    n = mib << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=device)
    d_out = torch.empty(n, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def run(h2d, d2h):
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(device)
        return n * reps / (time.perf_counter() - t0) / 1e9

    run(True, True)
    return {"h2d_gbs": run(True, False), "d2h_gbs": run(False, True), "both_gbs_each_direction": run(True, True)}


if __name__ == "__main__":
    rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(rank)
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    print(json.dumps({"rank": rank, **probe(torch.device("cuda", rank), mib)}), flush=True)
