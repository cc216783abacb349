"""Data-parallel plumbing: one process per GPU, torch.distributed for rendezvous, ONE NCCL
all-reduce of the flat gradient buffer per learner step (inside the CUDA graph).

The reference has no collective (its multi-worker mode is Hogwild over shared CPU memory,
main.py:394-405); this is the synchronous-DP layout of SURVEY.md section 8e: replay storage and
trees are sharded by rank (each rank owns its own ring / trees / max_priority), every rank
samples B_local rows from its shard, gradients are summed over ranks with 1/(B_local*world)
folded into the loss-gradient kernel, and the identical fused Adam runs on every rank so the
replicas stay bit-identical.
"""
import ctypes as C
import os

import torch


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n items owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload, nbytes, src=0, device="cpu"):
    """Broadcast a fixed-size byte string from `src` over the default process group
    (works on gloo with CPU tensors and on NCCL with CUDA tensors)."""
    import torch.distributed as dist
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.tensor(list(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


class Comm(object):
    """NCCL communicator owned by libd4pg_sm100.so (d4pg_comm_*), bootstrapped through the
    torch.distributed default group."""

    def __init__(self, rank=None, world_size=None, device=None):
        import torch.distributed as dist
        from . import _lib
        _lib.require_cuda()
        self.rank = dist.get_rank() if rank is None else rank
        self.world_size = dist.get_world_size() if world_size is None else world_size
        uid = (C.c_uint8 * 128)()
        if self.rank == 0:
            _lib.check(_lib.lib().d4pg_comm_unique_id(uid), "d4pg_comm_unique_id")
        dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
        raw = broadcast_bytes(bytes(uid), 128, src=0, device=dev)
        uid2 = (C.c_uint8 * 128).from_buffer_copy(raw)
        h = C.c_void_p()
        _lib.check(_lib.lib().d4pg_comm_create(uid2, self.rank, self.world_size, C.byref(h)), "d4pg_comm_create")
        self.handle = h

    def setup_peers(self, n_floats):
        """Fused all-reduce over peer memory (one node): allocate this rank's exchange block, gather the CUDA IPC
        handles of all ranks, map them.  Collective (every rank must call it); falls back to the NCCL all-reduce --
        on every rank -- if any rank cannot export or map (D4PG_COMM_PEER=0 disables it)."""
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        if self.world_size <= 1 or self.world_size > 8 or L.d4pg_comm_peer_ready(self.handle):
            return bool(L.d4pg_comm_peer_ready(self.handle))
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        ok = os.environ.get("D4PG_COMM_PEER", "1") != "0"
        mine = (C.c_uint8 * 64)()
        if ok and L.d4pg_comm_peer_alloc(self.handle, int(n_floats), mine) != 0:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            return False
        t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=dev)
        parts = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(self.world_size)]
        dist.all_gather(parts, t)
        blob = b"".join(bytes(p.cpu().tolist()) for p in parts)
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        ok = L.d4pg_comm_peer_open(self.handle, arr) == 0
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:                  # e.g. no CUDA IPC between the ranks' containers: NCCL path on every rank
            L.d4pg_comm_peer_disable(self.handle)
            return False
        return True

    def allreduce_sum_(self, flat):
        from . import _lib
        _lib.check(_lib.lib().d4pg_comm_allreduce_sum(self.handle, _lib.ptr(flat), flat.numel(), _lib.stream_ptr()),
                   "d4pg_comm_allreduce_sum")
        return flat

    def close(self):
        from . import _lib
        if self.handle is not None:
            _lib.lib().d4pg_comm_destroy(self.handle)
            self.handle = None
