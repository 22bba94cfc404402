"""Worker of test_gather_over_gloo_world_size_2 (launched by torch.distributed.run, gloo backend)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    n_items = int(sys.argv[1])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    sh = entry.load_package().sharding
    start, stop = sh.shard_range(n_items, rank, world)
    h, w = 4, 8
    full = torch.arange(n_items * h * w * 2, dtype=torch.float32).reshape(n_items, h, w, 2)
    out = sh.gather_erp_flow(full[start:stop].clone(), n_items, dst=0)
    if rank == 0:
        assert out is not None and torch.equal(out, full), "gathered tensor differs"
        print("GATHER_OK")
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
