"""Worker of test_peer_store_gather (launched by torch.distributed.run): every rank stitches its shard straight into
rank 0's CUDA-IPC-mapped buffer (sharding.PeerGatherBuffer).  The ranks may share one GPU (the 1-GPU test box): the
rendezvous then runs over gloo, which only carries the 64-byte IPC handle; with one GPU per rank it is the NCCL path of
bench.py.  No kernel waits on another rank's kernel, so time slicing on one device is safe."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    n_items, mode = int(sys.argv[1]), sys.argv[2]
    n_gpus = torch.cuda.device_count()
    rank_env = int(os.environ["RANK"])
    device = torch.device("cuda", rank_env % n_gpus)
    torch.cuda.set_device(device)
    dist.init_process_group("nccl" if n_gpus >= int(os.environ["WORLD_SIZE"]) else "gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = entry.load_package()
    sh, ops, lib = pkg.sharding, pkg._ops, pkg._lib
    erp_h, wt, pad = 128, 64, 0.3
    plan = pkg._geometry.get_plan(pad, erp_h, wt, device)
    gen = torch.Generator(device=device)
    gen.manual_seed(7)                       # every rank draws the whole batch, then keeps its shard
    src = torch.randint(0, 256, (n_items, erp_h, 2 * erp_h, 3), dtype=torch.uint8, device=device, generator=gen)
    tar = torch.randint(0, 256, (n_items, erp_h, 2 * erp_h, 3), dtype=torch.uint8, device=device, generator=gen)
    flows = (torch.rand((n_items, plan.tangent_pixels, 2), device=device, generator=gen) * 6.0 - 3.0).contiguous()
    start, stop = sh.shard_range(n_items, rank, world)
    buf = sh.PeerGatherBuffer(n_items, erp_h, 2 * erp_h, dst=0, device=device)
    assert (buf.start, buf.stop) == (start, stop)
    local = torch.empty((stop - start, erp_h, 2 * erp_h, 2), dtype=torch.float32, device=device)
    for blend in (lib.BLEND_STRAIGHTFORWARD, lib.BLEND_NORMWARP):
        args = (plan, flows[start:stop].contiguous(), src[start:stop].contiguous(), tar[start:stop].contiguous(), blend, True)
        if mode == "fused":                  # the blend kernel's stores land in rank 0's memory
            ops.ico2erp_flow_out(*args, buf.slice_ptr())
        else:                                # local result, then one asynchronous copy into the peer-mapped slice
            ops.ico2erp_flow_out(*args, local)
            lib.check(lib.load().tif_ipc_copy_async(buf.slice_ptr(), local.data_ptr(), local.numel() * 4,
                                                    torch.cuda.current_stream().cuda_stream), "tif_ipc_copy_async")
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            want = torch.empty((n_items, erp_h, 2 * erp_h, 2), dtype=torch.float32, device=device)
            ops.ico2erp_flow_out(plan, flows, src, tar, blend, True, want)
            torch.cuda.synchronize()
            assert torch.equal(buf.tensor(), want), "collected flows differ from a single-rank stitch (blend %d)" % blend
        dist.barrier()
    buf.close()
    if rank == 0:
        print("PEER_GATHER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
