"""Multi-rank worker (launched by torch.distributed.run): sharded decimation of one synthetic stream.

mode "gpu": each rank generates its chunk on its GPU, runs sdr_decimate_sharded (interior + NCCL halo + boundary) and
            the position-weighted checksums of the shards must add up to the checksum of the single-GPU result;
            boundary outputs are also compared with the CPU oracle.
mode "cpu": the same plan driven on the CPU with gloo send/recv for the halo and the oracle port as the kernel --
            covers the host-side sharding logic without a GPU.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402

TAPS, FACTOR = 128, 8


def main():
    mode, log2n = sys.argv[1], int(sys.argv[2])
    n = (1 << log2n) + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    taps = synth.windowed_sinc_taps(TAPS, 1 / 16)
    from sdr_b200.multigpu import shard_plan
    plan = shard_plan(n, TAPS, FACTOR, world, rank)
    total_out = (n - TAPS) // FACTOR + 1

    if mode == "cpu":
        import oracle
        port = oracle.port()
        x = synth.noise_complex(plan.in_count, first=plan.in_begin)
        # halo exchange: my first `left.halo` samples go to rank-1, I receive plan.halo samples from rank+1
        reqs = []
        if rank > 0:
            left = shard_plan(n, TAPS, FACTOR, world, rank - 1)
            if left.halo:
                reqs.append(dist.isend(torch.from_numpy(x[:left.halo].view(np.float32).copy()), rank - 1))
        halo = torch.zeros(2 * plan.halo, dtype=torch.float32)
        if plan.halo:
            reqs.append(dist.irecv(halo, rank + 1))
        for r in reqs:
            r.wait()
        xs = np.concatenate([x, halo.numpy().view(np.complex64)])
        local0 = plan.out_begin * FACTOR - plan.in_begin
        y = port.decimate(oracle.V_AVX, plan.out_count, FACTOR, np.repeat(taps, 2), xs[local0:], True)
        part = synth.checksum32(y.view(np.uint32), first=2 * plan.out_begin)
        parts = [None] * world
        dist.all_gather_object(parts, (part, plan.out_count))
        if rank == 0:
            whole = port.decimate(oracle.V_AVX, total_out, FACTOR, np.repeat(taps, 2), synth.noise_complex(n), True)
            want = synth.checksum32(whole.view(np.uint32))
            got = sum(p[0] for p in parts) & 0xffffffffffffffff
            assert sum(p[1] for p in parts) == total_out, (parts, total_out)
            assert got == want, (hex(got), hex(want))
            print("MG_OK cpu", world, n)
        dist.destroy_process_group()
        return

    import sdr_b200
    from sdr_b200 import _lib as L
    from sdr_b200 import multigpu
    ctx = sdr_b200.Context(int(os.environ.get("LOCAL_RANK", rank)))
    uid = multigpu.unique_id() if rank == 0 else bytes(L.COMM_ID_BYTES)
    t = torch.tensor(list(uid), dtype=torch.uint8)
    dist.broadcast(t, 0)
    comm = multigpu.Comm(ctx, bytes(t.tolist()), world, rank)
    dec = sdr_b200.cudaDecimatorC(FACTOR, taps, ctx=ctx, sizeMultiple=4)
    d_in = ctx.alloc(8 * plan.in_count + 64)
    d_out = ctx.alloc(8 * max(plan.out_count, 1) + 64)
    ctx.synth_noise(d_in, 2 * plan.in_count, first_float=2 * plan.in_begin)
    if mode == "gpu-peer":
        ctx.sync()
        dist.barrier()                      # every rank's chunk is complete before anyone reads a neighbour's
        comm.share_chunks(d_in.ptr)
        assert comm.peer_halo_active(d_in.ptr)
    for _ in range(3):   # repeated passes reuse the halo buffer and events
        multigpu.decimate_sharded(dec, comm, plan, d_in.ptr, d_out.ptr)
    ctx.sync()
    part = ctx.checksum32(d_out, 2 * plan.out_count, first_word=2 * plan.out_begin)
    # boundary outputs (the ones that needed the neighbour's samples) against the CPU oracle
    ok = True
    nb = plan.out_count - plan.out_interior
    if nb > 0:
        import oracle
        first = plan.out_begin + plan.out_interior
        xs = synth.noise_complex(nb * FACTOR + TAPS, first=first * FACTOR)
        want = oracle.port().decimate(oracle.V_AVX, nb, FACTOR, np.repeat(taps, 2), xs, True)
        got = d_out.to_host(np.complex64, nb, offset_bytes=8 * plan.out_interior)
        scale = np.maximum(np.abs(want), np.sqrt(np.mean(np.abs(want) ** 2)))
        ok = bool(np.all(np.abs(got - want) <= 1e-5 * scale))
    parts = [None] * world
    dist.all_gather_object(parts, (part, plan.out_count, ok, nb))
    if rank == 0:
        assert all(p[2] for p in parts), parts
        assert sum(p[1] for p in parts) == total_out
        assert sum(p[3] for p in parts) == (world - 1) * (TAPS - FACTOR) // FACTOR, parts
        # single-GPU result of the whole stream on rank 0's device
        x = ctx.alloc(8 * n + 64)
        y = ctx.alloc(8 * total_out + 64)
        ctx.synth_noise(x, 2 * n)
        L.check(L.lib.sdr_decimate_stream(dec.handle, x.ptr, n, y.ptr, total_out))
        want = ctx.checksum32(y, 2 * total_out)
        got = sum(p[0] for p in parts) & 0xffffffffffffffff
        assert got == want, (hex(got), hex(want))
        print("MG_OK", mode, world, n)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
