// ring_common.cuh -- PTX helpers shared by the persistent shared-memory-ring kernels (kernels_fast.cu, kernels_real.cu):
// packed FP32 FMA, mbarrier, TMA bulk copy, and the contiguous-slot ring used by the real-data kernels.
#pragma once
#include "common.cuh"

namespace sdr {

typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 dup2(float v) {
    u64 d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(v));
    return d;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}


// Generation guard.  An mbarrier wait only carries ONE parity bit, so a warp that runs two generations ahead of a slot
// (it may: different warps consume successive generations of the same slot) would see the phase "two back" as
// completed and read stale data.  The filler therefore publishes the generation it has armed in a plain
// shared-memory word, and a consumer first spins until that word says its generation has been armed -- after which
// the parity wait is unambiguous.
// No fence between arming the barrier and publishing: both are shared-memory operations of the SAME thread to the same
// SM's shared memory, which the LSU performs in program order; a CTA-scope fence here would also wait for the lane's
// in-flight output stores (measured: -10 % on the decimator).
__device__ __forceinline__ void gen_publish(uint32_t addr, int value) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(value) : "memory");
}
__device__ __forceinline__ int gen_read(uint32_t addr) {
    int v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// Wait until generation `gen` (1-based) of a slot has landed.  The published generation is read BEFORE the parity
// wait (its latency hides behind the wait; a polling loop on it would steal issue slots from the warp computing on the
// same SMSP, and a read after the wait would sit exposed in front of the first data load): if the slot was already
// armed for `gen` the parity wait is unambiguous.  Only when it was not -- the rare case in which the wait may have
// matched the phase two generations back -- does the warp back off until it is armed and wait again.
// GUARD = false when the number of slots is a multiple of the number of warps: then every generation of a slot is
// consumed by the same warp, in order, and the one-bit parity cannot alias.
// the rare path lives out of line so that it cannot disturb the register allocation / scheduling of the hot loop
static __device__ __noinline__ void slot_wait_slow(uint32_t bar, uint32_t gen_addr, int gen) {
    while (gen_read(gen_addr) < gen) __nanosleep(64);
    mbar_wait(bar, (gen - 1) & 1);
}
template <bool GUARD>
__device__ __forceinline__ void slot_wait(uint32_t bar, uint32_t gen_addr, int gen) {
    if (!GUARD) { mbar_wait(bar, (gen - 1) & 1); return; }
    const int armed = gen_read(gen_addr);
    mbar_wait(bar, (gen - 1) & 1);
    if (armed < gen) slot_wait_slow(bar, gen_addr, gen);
}

// Ring of NS contiguous slots of SLOT_BYTES in shared memory, filled by ONE TMA bulk copy per slot (issued by lane 0
// of the warp that owns the refill), plus a mirror of the first HALO_BYTES of slot 0 behind the last slot so that a
// window running off the end of slot NS-1 reads on linearly.  Local slot index u lives in ring slot u % NS; slot
// index `cnt` is a halo-only fill (HALO_BYTES).  full[s]: 1 arrival + tx bytes; empty[s]: 2 arrivals (the warp that
// computed the slot and the warp that read its head as halo).
template <int SLOT_BYTES, int HALO_BYTES, int NS>
struct ContigRing {
    static_assert(SLOT_BYTES % 16 == 0 && HALO_BYTES % 16 == 0 && HALO_BYTES <= SLOT_BYTES, "TMA bulk copies move 16-byte units");
    static constexpr int RING_BYTES = NS * SLOT_BYTES + HALO_BYTES;
    static constexpr int BAR_OFFSET = ((RING_BYTES + 127) / 128) * 128;
    static constexpr int SMEM_BYTES = BAR_OFFSET + 2 * NS * 8 + NS * 4 + 128;
    uint32_t ring, bar_full, bar_empty, gen_armed;
    const unsigned char *src0;   // global address of local slot 0
    int cnt;

    __device__ __forceinline__ void init(unsigned char *smem, const unsigned char *src, int n_slots) {
        ring = smem_u32(smem); bar_full = ring + BAR_OFFSET; bar_empty = bar_full + NS * 8; gen_armed = bar_empty + NS * 8;
        src0 = src; cnt = n_slots;
        if (threadIdx.x == 0) {
            for (int s = 0; s < NS; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 2);
                                           asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(gen_armed + 4 * s), "r"(0) : "memory"); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
    }
    __device__ __forceinline__ void issue_fill(int u, int lane) {
        if (lane != 0) return;
        const int slot = u % NS;
        const uint32_t bytes = (u == cnt) ? HALO_BYTES : SLOT_BYTES;
        const uint32_t bar = bar_full + 8 * slot;
        const unsigned char *src = src0 + (long long)u * SLOT_BYTES;
        mbar_expect_tx(bar, bytes + (slot == 0 ? HALO_BYTES : 0));
        bulk_g2s(ring + slot * SLOT_BYTES, src, bytes, bar);
        if (slot == 0) bulk_g2s(ring + NS * SLOT_BYTES, src, HALO_BYTES, bar);
        gen_publish(gen_armed + 4 * slot, u / NS + 1);
    }
    __device__ __forceinline__ void prologue(int warp, int lane, int n_warps) {
        for (int u = warp; u < NS && u <= cnt; u += n_warps) issue_fill(u, lane);
    }
    __device__ __forceinline__ void wait_slot(int u) {
        slot_wait<(NS % 8) != 0>(bar_full + 8 * (u % NS), gen_armed + 4 * (u % NS), u / NS + 1);
        slot_wait<(NS % 8) != 0>(bar_full + 8 * ((u + 1) % NS), gen_armed + 4 * ((u + 1) % NS), (u + 1) / NS + 1);
    }
    // after the warp has finished reading slot u (and the head of slot u+1)
    __device__ __forceinline__ void release_and_refill(int u, int lane) {
        __syncwarp();
        const int slot = u % NS;
        if (lane == 0) {
            mbar_arrive(bar_empty + 8 * slot);
            if (u == 0) mbar_arrive(bar_empty + 8 * slot);   // slot 0 has no predecessor using it as halo
            mbar_arrive(bar_empty + 8 * ((u + 1) % NS));
        }
        if (u + NS <= cnt) {
            mbar_wait(bar_empty + 8 * slot, (u / NS) & 1);
            issue_fill(u + NS, lane);
        }
    }
};

}  // namespace sdr
