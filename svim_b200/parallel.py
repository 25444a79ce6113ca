"""One process per GPU: shard bookkeeping and rendezvous around the two NCCL all-gather-v
calls of the C ABI (svimgpu_exchange_signatures, svimgpu_cluster_sharded).

`torch.distributed` is used only as the bootstrap channel (ranks, NCCL unique id, record
counts); the data path never goes through it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .records import AlignmentBatch


def shard_ranges(n_cigar: np.ndarray, world: int):
    """Contiguous record ranges [lo, hi) per rank, balanced by CIGAR bytes (the COLLECT cost).
    Contiguity keeps the reference's emission order = concatenation in rank order."""
    w = np.concatenate([[0], np.cumsum(n_cigar.astype(np.int64) + 11)])   # +11 words ~ the fixed row
    total = int(w[-1])
    cuts = [int(np.searchsorted(w, total * r / world, side="left")) for r in range(world)] + [len(n_cigar)]
    cuts[0] = 0
    for r in range(1, world + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_batch(batch: AlignmentBatch, rank: int, world: int) -> AlignmentBatch:
    """This rank's record range with its own blobs (offsets rebased): only the shard is uploaded and held in HBM."""
    lo, hi = shard_ranges(batch.n_cigar, world)[rank]
    return batch.compact_slice(lo, hi)


def exchange_layout(n_local: int):
    """(aln_base, total) from an all-gather of the per-rank record counts."""
    import torch.distributed as dist
    world = dist.get_world_size()
    sizes = [None] * world
    dist.all_gather_object(sizes, int(n_local))
    rank = dist.get_rank()
    return int(sum(sizes[:rank])), int(sum(sizes)), sizes


def init_comm(ctx: "_lib.Context"):
    """Create the NCCL communicator of `ctx` from rank 0's unique id."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    ids = [None]
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        rc = ctx.lib.svimgpu_nccl_unique_id(buf)
        if rc != 0:
            raise _lib.SvimGpuError(rc, "ncclGetUniqueId failed")
        ids = [bytes(buf)]
    dist.broadcast_object_list(ids, src=0)
    idb = (C.c_ubyte * 128).from_buffer_copy(ids[0])
    ctx._check(ctx.lib.svimgpu_comm_init(ctx.h, world, rank, idb))


def collect_and_cluster(ctx: "_lib.Context", aln_base: int, which: int = 0):
    """collect (local shard, already uploaded) -> exchange -> sharded cluster; every rank returns the full result.
    The inserted sequences stay on the rank that collected them (include/svimgpu.h, svimgpu_exchange_signatures): barrier between
    a `fetch_signatures` of the gathered lists and the next collect of any rank."""
    cst = ctx.collect()
    xst = _lib.CollectStats()
    ctx._check(ctx.lib.svimgpu_exchange_signatures(ctx.h, aln_base, C.byref(xst)))
    ctx.use_collected(which)
    clst, clusters, members = ctx.cluster(sharded=True)
    return cst, xst, clst, clusters, members
