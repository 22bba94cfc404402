"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed).

Frame pairs are independent units, so the hot path shards on the batch axis with no data-path
collective.  The only exchange is the optional collection of the stitched ERP
flows on one rank: a single NCCL gather over NVLink (gloo on CPU for the tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced [start, stop) of `n_items` for `rank`; the first n % world ranks get one extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank %r / world size %r" % (rank, world_size))
    base, extra = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world_size):
    return [shard_range(n_items, r, world_size)[1] - shard_range(n_items, r, world_size)[0] for r in range(world_size)]


def gather_erp_flow(local_flow, n_items, dst=0, group=None):
    """Collect per-rank stitched flows [b_r, H, W, 2] into [n_items, H, W, 2] on rank `dst`.

    Shards follow shard_range().  Equal shards use one dist.gather; ragged shards are padded to the
    largest shard first.  Returns the full tensor on `dst`, None elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_items, world)
    if local_flow.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d items, expected %d" % (rank, local_flow.shape[0], sizes[rank]))
    biggest = max(sizes)
    send = local_flow.contiguous()
    if send.shape[0] != biggest:
        pad = torch.zeros((biggest - send.shape[0],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        send = torch.cat((send, pad), 0)
    if rank == dst:
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, recv, dst=dst, group=group)
        return torch.cat([recv[r][:sizes[r]] for r in range(world)], 0)
    dist.gather(send, None, dst=dst, group=group)
    return None
