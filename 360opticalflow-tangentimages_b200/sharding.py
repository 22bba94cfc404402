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


class PeerGatherBuffer:
    """The collected [n_items, H, W, 2] float32 flow field, resident on rank `dst` and mapped into every other rank's
    address space through CUDA IPC (tif_ipc_*): a rank stitches its shard straight into `slice_ptr()` -- the blend
    kernel's stores cross NVLink / NVSwitch themselves, so the gather costs no extra kernel, copy or receive step.

    Collective constructor (every rank of `group` must call it).  `torch.distributed` only carries the 64-byte handle.
    """

    def __init__(self, n_items, erp_h, erp_w, dst=0, group=None, device=None):
        import ctypes as C

        from . import _lib
        self.lib = _lib.load()
        self.group, self.dst = group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_items, self.erp_h, self.erp_w = int(n_items), int(erp_h), int(erp_w)
        self.item_bytes = self.erp_h * self.erp_w * 2 * 4
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.base = C.c_void_p()
        self._owner = self.rank == dst
        box = [None]
        with torch.cuda.device(self.device):
            if self._owner:
                _lib.check(self.lib.tif_ipc_alloc(C.byref(self.base), self.n_items * self.item_bytes), "tif_ipc_alloc")
                handle = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
                _lib.check(self.lib.tif_ipc_export(self.base, handle), "tif_ipc_export")
                box[0] = handle.raw
            dist.broadcast_object_list(box, src=dst, group=group)
            if not self._owner:
                _lib.check(self.lib.tif_ipc_open(box[0], C.byref(self.base)), "tif_ipc_open")
        self.start, self.stop = shard_range(self.n_items, self.rank, self.world)

    def slice_ptr(self, item=None):
        """Device address (valid on this rank) of item `item` of the collected buffer (default: this rank's first)."""
        item = self.start if item is None else int(item)
        return int(self.base.value) + item * self.item_bytes

    def tensor(self):
        """The collected buffer as a torch tensor (rank `dst` only; zero copy)."""
        if not self._owner:
            return None

        class _Mem:
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (self.n_items, self.erp_h, self.erp_w, 2), "typestr": "<f4",
                                      "data": (int(self.base.value), False), "version": 3, "strides": None}
        self._keep = m
        return torch.as_tensor(m, device=self.device)

    def close(self):
        if self.base is not None and self.base.value:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
                dist.barrier(group=self.group)       # nobody may still be writing when the owner frees
                if self._owner:
                    self.lib.tif_ipc_free(self.base)
                else:
                    self.lib.tif_ipc_close(self.base)
            self.base = None
