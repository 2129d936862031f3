"""Host-side helpers for the two multi-GPU modes (one process per GPU, torch.distributed for the plumbing).

chain-parallel  X, y replicated; rank r runs chains [r*C, (r+1)*C) with disjoint Philox substreams
                (Engine(chain_offset=r*C)); no collective in the data path.
row-sharded     rank r holds rows shard_rows(n, world, r) of X, y, eta; every pass the C*K per-candidate
                partial sums are all-gathered and summed in RANK ORDER, so the totals -- and therefore every
                accept/reject branch -- are bit-identical on all ranks.  In production the exchange runs
                inside the library over NCCL (Engine.comm_init_nccl); `ordered_sum_exchange` is the same
                arithmetic on torch tensors (any backend), used by the tests and as a reference.
"""
import ctypes as C
import numpy as np
from . import _lib as L


def shard_rows(n, world, rank):
    """Contiguous row block of `rank`: boundaries are even (128-bit alignment of the fp64 pairs a lane loads)
    and the blocks differ by at most 2 rows."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    pairs = (n + 1) // 2
    lo = 2 * ((pairs * rank) // world)
    hi = min(n, 2 * ((pairs * (rank + 1)) // world))
    return lo, hi


def ordered_sum(parts):
    """Sum of per-rank vectors in rank order (what rank_sum_kernel does on the device)."""
    out = np.zeros_like(np.asarray(parts[0], dtype=np.float64))
    for p in parts:
        out = out + np.asarray(p, dtype=np.float64)
    return out


def ordered_sum_exchange(t, group=None):
    """In-place cross-rank sum of tensor `t` with a rank-ordered reduction: all_gather + sequential adds.
    Unlike all_reduce this gives the same bits on every rank for any backend and any algorithm."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous(), group=group)
    acc = torch.zeros_like(t)
    for p in parts:
        acc = acc + p
    t.copy_(acc)
    return t


class DeviceBuffer:
    """Exposes a raw device pointer as a CUDA array so torch can wrap it without copying."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    L.check(L.load().cgg_nccl_unique_id(buf))
    return buf.raw


def init_nccl(engine, rank, world, group=None):
    """Creates the engine's NCCL communicator: rank 0 makes the id, torch.distributed carries the 128 bytes."""
    import torch.distributed as dist
    obj = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0, group=group)
    engine.comm_init_nccl(rank, world, obj[0])


def init_p2p(engine, rank, world, group=None):
    """Connects the peer mailboxes of a row-sharded engine on the persistent driver: every rank exports the CUDA IPC handle
    of its mailbox, torch.distributed carries the 64 bytes, every rank maps the others' (NVLink peer memory)."""
    import torch.distributed as dist
    _, handle = engine.p2p_mailbox(world)
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    engine.p2p_connect(rank, world, ipc_handles=handles)
