"""One process per GPU: row sharding over the slow (spin-down) index and the plumbing around the NCCL communicator
the engine owns.  torch.distributed is used only to move the 128-byte NCCL unique id and to take max-over-ranks timings.
"""
import ctypes as C
import os

from . import _lib
from ._lib import check


def shard_range(n, rank, nranks):
    """[first, first+count) of n items owned by `rank` (the engine's own rule, lpp_shard_range)."""
    f, c = C.c_uint64(), C.c_uint64()
    check(_lib.lib().lpp_shard_range(n, rank, nranks, C.byref(f), C.byref(c)))
    return f.value, c.value


def local_rows(n_up, n_down, rank, nranks):
    """Rows of the product basis (index = iup + idn*n_up) owned by `rank`: whole up-segments of a contiguous idn range."""
    f, c = shard_range(n_down, rank, nranks)
    return f * n_up, c * n_up


def broadcast_unique_id(dist, make_id, src=0):
    """rank `src` creates the id (engine.comm_unique_id on a GPU box), everybody receives the same 128 bytes."""
    ids = [make_id() if dist.get_rank() == src else None]
    dist.broadcast_object_list(ids, src=src)
    return ids[0]


def attach(engine, dist):
    """Give a row-sharded InternalProductCuda its NCCL communicator."""
    from .engine import comm_unique_id
    engine.comm_init(broadcast_unique_id(dist, comm_unique_id))
    engine._dist = dist                       # new-sector handles (engine.sector) repeat the peer-memory exchange through it
    if os.environ.get("LPP_P2P", "1") != "0":
        mine = engine.p2p_export()
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, mine)
        if all(p is not None for p in parts):
            engine.p2p_import(b"".join(parts))
            engine._has_p2p = True
    return engine
