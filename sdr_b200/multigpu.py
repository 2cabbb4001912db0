"""Chunk-sharding of a flat stream over the GPUs of one box with an NCCL halo exchange."""
import ctypes as C

from . import _lib as L


def shard_plan(n_samples, taps, factor, world, rank) -> L.ShardPlan:
    """pure host arithmetic: which input chunk and which outputs rank `rank` owns, and the halo it needs"""
    p = L.ShardPlan()
    L.check(L.lib.sdr_shard_plan(n_samples, taps, factor, world, rank, C.byref(p)))
    return p


def unique_id() -> bytes:
    buf = (C.c_ubyte * L.COMM_ID_BYTES)()
    L.check(L.lib.sdr_comm_unique_id(buf))
    return bytes(buf)


class Comm:
    """sdr_comm_t: an NCCL communicator bound to a Context"""

    def __init__(self, ctx, uid: bytes, world: int, rank: int):
        h = C.c_void_p()
        buf = (C.c_ubyte * L.COMM_ID_BYTES).from_buffer_copy(uid)
        L.check(L.lib.sdr_comm_create(ctx.h, buf, world, rank, C.byref(h)))
        self.h, self.ctx, self.world, self.rank = h, ctx, world, rank

    def share_chunks(self, d_chunk_base):
        """collective: map the right neighbour's chunk (CUDA IPC) so the halo is read in place over NVLink"""
        L.check(L.lib.sdr_comm_share_chunks(self.h, d_chunk_base))

    def barrier(self):
        """in-stream rendezvous (4-byte ncclAllReduce on the ctx stream); enqueue-only"""
        L.check(L.lib.sdr_comm_barrier(self.h))

    def peer_halo_active(self, d_in) -> bool:
        return bool(L.lib.sdr_comm_peer_halo_active(self.h, d_in))

    def close(self):
        if self.h:
            L.lib.sdr_comm_destroy(self.h)
            self.h = None


def decimate_sharded(decimator, comm, plan, d_in, d_out):
    """one pass of the sharded decimator (interior, halo exchange, boundary); enqueue-only"""
    L.check(L.lib.sdr_decimate_sharded(decimator.handle, comm.h if comm else None, C.byref(plan), d_in, d_out))
