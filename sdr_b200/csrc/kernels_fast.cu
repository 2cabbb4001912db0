// kernels_fast.cu -- tuned sm_100a kernels for the headline shapes.
//
// dec_c_ring<T, D, R>: complex data x real taps, decimate by D (reference decimate.c:73-113 `decimateRC` /
// `decimateAVXRC`, reached from fastDecimatorC, Filter.hs:352-356).  y[m] = sum_k c[k] x[m*D + k].
//
// Design (DESIGN.md section "decimator kernel"):
//   * persistent: one CTA per SM, 8 warps, each CTA walks a contiguous range of SUB-TILES (32*R outputs each).
//   * the input stream is staged ONCE into a shared-memory ring of NS slots (one sub-tile = 32 lane segments of
//     R*D complex samples) by TMA bulk copies (cp.async.bulk -> UBLKCP), completion on mbarriers; a slot is refilled
//     by the warp that consumed it as soon as it and the neighbour that used its head as halo have arrived on the
//     slot's `empty` barrier.  ~4 slots (64 KB) are in flight per SM.
//   * lane l of a warp owns R consecutive outputs; its window is R + T/D - 1 decimation blocks (D samples) starting at
//     its own segment.  Segments are padded by 16 B so the lanes' LDS.128 hit distinct bank groups.
//   * taps live in T registers; the inner product is fully unrolled FFMA2 (fma.rn.f32x2: re and im of one sample
//     in one instruction, tap broadcast), accumulation in increasing tap order -- the same order as the generic
//     kernel, so the ragged tail finished by launch_fir_generic is computed identically.
//   * no tensor cores (1-D dot products); the kernel needs 2*T/D = 32 FMA per 9 algorithmic bytes, i.e. it is
//     balanced between HBM and the FP32 pipe (see DESIGN.md roofline).
// This unit: the dispatcher (launch_dec_fast) and the persistent consumer; the ring kernel template is in dec_ring.cuh.
#include "dec_ring.cuh"

namespace sdr {

// true when launch_dec_fast would produce ALL `num` outputs in one launch (no generic tail to fork): same conditions as
// the covering mode of launch_ring
bool dec_fast_will_cover(bool cplx, int taps_stored, int D, Seg2 seg, long long num) {
    if (num <= 0 || taps_stored > 256 || (D != 4 && D != 8 && D != 16)) return false;
    const int epc = cplx ? 2 : 4;
    return (((uintptr_t)seg.a) & 15) == 0 && (seg.na % epc) == 0 && (seg.nb == 0 || (((uintptr_t)seg.b) & 15) == 0) &&
           ((seg.na + seg.nb) % epc) == 0;
}

// x = seg.a ++ seg.b.  taps_stored = the record's tap count (d_taps is zero-padded to at least 128 floats).  Dispatch only:
// the instantiations live in kernels_fast_c.cu (complex), kernels_fast_r.cu (real) and kernels_fast_p.cu (launch-parameter forms).
int launch_dec_fast(Ctx *c, bool cplx, int taps_stored, int D, const float *d_taps, Seg2 seg, void *d_out, long long num,
                    long long *done, const char **name, const float *h_taps) {
    *done = 0;
    *name = cplx ? "fir_direct" : "fir_tile";
    if ((((uintptr_t)seg.a) & 15) != 0) return SDR_OK;   // TMA bulk copies need a 16-byte aligned source; any output alignment
    if (D != 4 && D != 8 && D != 16) return SDR_OK;
    const char *label = nullptr;
    if (taps_stored > 128) {
        // 129..256 taps: the taps cannot live in registers -- they travel as launch parameters and reach every FFMA / FFMA2
        // as a uniform-register operand (h_taps: the record's host copy, zero-padded here).  78 registers instead of 200.
        // FP32-pipe bound at decimation 4 and 8 (128 / 64 FMA per complex input sample), HBM-bound at 16.
        if (!(taps_stored <= 256 && h_taps)) return SDR_OK;
        float padded[256] = {0.0f};
        for (int k = 0; k < taps_stored; k++) padded[k] = h_taps[k];
        SDR_TRY(launch_ring_param(c, cplx, 256, D, d_taps, padded, seg, d_out, num, done, &label));
        if (label) *name = label;
        return SDR_OK;
    }
    const int T = taps_stored <= 32 ? 32 : taps_stored <= 64 ? 64 : 128;
    // 128 taps, real data at decimation 4 / 8 and complex data at decimation 4: taps as launch parameters (uniform-register
    // operands), which frees the registers for 16 warps at 8 outputs per lane -- real 588 / 1055 Gsamples/s against 506 / 969
    // for the register-tap form, complex 432 against 380 (the register form stays for 64 taps, where it is the faster one:
    // 824 / 1486 against 760 / 1397, and for complex decimation 8 / 16, which sit at the HBM roofline).
    // SDR_B200_DEC_TP=0 switches it off, =2 also takes the real 64-tap shapes (measurement knob).
    static const int tp_mode = getenv("SDR_B200_DEC_TP") ? atoi(getenv("SDR_B200_DEC_TP")) : 1;
    if (tp_mode && h_taps && ((!cplx && (D == 4 || D == 8) && (T == 128 || (tp_mode == 2 && T == 64))) || (cplx && D == 4 && T == 128))) {
        float padded[128] = {0.0f};
        for (int k = 0; k < taps_stored; k++) padded[k] = h_taps[k];
        SDR_TRY(launch_ring_param(c, cplx, T, D, d_taps, padded, seg, d_out, num, done, &label));
        if (label) { *name = label; return SDR_OK; }
    }
    if (cplx) SDR_TRY(launch_ring_complex(c, T, D, d_taps, seg, d_out, num, done, &label));
    else      SDR_TRY(launch_ring_real(c, T, D, d_taps, seg, d_out, num, done, &label));
    if (label) *name = label;
    return SDR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent consumer: the ring decimator fed by a flag-published stream (the Pipes per-vector contract on the device
// without a launch per vector).  The host publishes how many bytes of a contiguous device-resident stream are valid in a
// page-locked control block the kernel polls; the kernel publishes the runs it has finished in the same block.
//
//   * work unit = a RUN of PB consecutive sub-tiles (PB * 32 * R outputs); run r belongs to CTA r % G.  A CTA's tiles
//     form one sequence t = q * (PB + 1) + j over its runs q: j < PB is a computed sub-tile, j == PB is the halo-only
//     slot behind the run (the head of the next run, which another CTA computes).  The shared-memory ring, the TMA fills
//     and the mbarrier protocol are those of k_dec_ring and keep flowing across runs: there is no per-run bubble while the
//     input is ahead.
//   * a fill is issued only once its whole run (halo included) has been published.  The warp that frees a slot refills it
//     if the target run is already available (no waiting); otherwise the warp that will CONSUME the tile issues the fill
//     itself once the run arrives -- a per-slot claim word (shared-memory CAS on the generation) makes exactly one of them
//     do it.  So a warp only ever waits for input it needs itself, and every published run is computed even if nothing
//     further is ever published.
//   * when the last sub-tile of a run has been stored (per-run counter in shared memory), the run's flag is set in host
//     memory behind a system-scope fence; the host then pops its outputs with ordinary copies on another stream.
//   * `closed` ends the session: runs not completely published by then are left to the ordinary launch path.  A warp that
//     waits for input longer than ~3 s sets `error` and leaves (the host never gets a stuck GPU).
struct PersistCtl {                      // lives in cudaHostAllocMapped memory; one cache line per writer
    volatile long long published_bytes;  // host -> device
    volatile int closed;                 // host -> device
    int pad0_[13];
    volatile int error;                  // device -> host (byte 64)
    int pad1_[15];
    volatile unsigned int done[1];       // device -> host (byte 128): done[r] = 1 when run r is complete
};
// what the relay CTA copies from host memory into device memory for the worker CTAs to poll: 1100 warps polling a host
// cache line over PCIe would both saturate the link with reads and slow the host's stores to that line to a crawl
struct PersistRelay { volatile long long avail; volatile int closed; volatile int error; };

template <bool CPLX, int T, int D, int R, int PB>
__global__ void __launch_bounds__(256, 1)
k_dec_ring_persist(const void *__restrict__ in, void *__restrict__ out, const float *__restrict__ taps, PersistCtl *ctl,
                   PersistRelay *relay, long long runs_total, int dbg_flags) {
    typedef RingCfg<CPLX, T, D, R> C;
    static_assert(CPLX, "the persistent consumer is instantiated for complex data");
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x - 1, cta = blockIdx.x;
    if (cta == G) {
        // the relay CTA (one thread): host memory -> device memory, until the session is closed.  `closed` is read before
        // the byte count and written after it, so a count seen together with closed == 1 is final.
        if (threadIdx.x != 0) return;
        unsigned long long t_last, t_now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_last));
        long long prev = -1;
        for (;;) {
            const int c = ctl->closed;
            __threadfence_system();                      // the byte count is read AFTER the flag
            const long long a = ctl->published_bytes;
            relay->avail = a;
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(&relay->closed), "r"(c) : "memory");
            if (c || relay->error) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
            if (a != prev) { prev = a; t_last = t_now; }
            else if (t_now - t_last > 3000000000ULL) { relay->error = 1; ctl->error = 1; relay->closed = 1; break; }   // idle for 3 s
        }
        return;
    }
    const bool vec_store = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    constexpr long long SUB_BYTES = 32LL * C::SEG_BYTES;
    constexpr int PERIOD = PB + 1;

    const uint32_t ring = smem_u32(smem);
    const uint32_t bar_full = ring + C::RING_BYTES + ((128 - C::RING_BYTES % 128) % 128);
    const uint32_t bar_empty = bar_full + C::NS * 8;
    const uint32_t gen_armed = bar_empty + C::NS * 8;
    __shared__ int s_claim[C::NS];       // last generation of each slot whose fill has been claimed
    __shared__ int s_run_cnt[8];         // warps that have stored their share of a run in flight (indexed by q % 8)
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NS; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 2);
                                          asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(gen_armed + 4 * s), "r"(0) : "memory");
                                          s_claim[s] = 0; }
        for (int i = 0; i < 8; i++) s_run_cnt[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const unsigned char *gin = reinterpret_cast<const unsigned char *>(in);
    long long avail = 0;      // bytes known to be published (warp-uniform cache of ctl->published_bytes)
    int closed = 0;
    auto run_of = [&](int t) -> long long { return (long long)(t / PERIOD) * G + cta; };
    auto need_of_run = [&](long long r) -> long long { return ((r + 1) * PB * 32 + C::HALO_SEGS) * (long long)C::SEG_BYTES; };
    // refresh the cache from host memory (one lane reads, all lanes get the values); `closed` is read FIRST so that a
    // byte count read after it is final
    auto refresh = [&]() {
        long long a = 0; int c = 0;
        if (lane == 0) {   // device memory (L2), written by the relay CTA; the acquire orders the count's load after the flag's
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(c) : "l"(&relay->closed) : "memory");
            a = relay->avail;
        }
        closed = __shfl_sync(0xffffffffu, c, 0);
        avail = __shfl_sync(0xffffffffu, a, 0);
    };
    auto have_run = [&](long long r, bool wait) -> bool {
        const long long need = need_of_run(r);
        if (avail >= need) return true;
        if (closed) return false;
        refresh();
        if (avail >= need || !wait) return avail >= need;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (avail < need) {
            if (closed) return false;
            __nanosleep(400);
            refresh();
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 4000000000ULL) { if (lane == 0) { relay->error = 1; ctl->error = 1; } closed = 1; return false; }
            if (relay->error) { closed = 1; return false; }
        }
        return true;
    };
    // fill tile t (all lanes): its run is known to be published
    auto issue_fill = [&](int t) {
        const int slot = t % C::NS, j = t % PERIOD;
        const int nseg = (j == PB) ? C::HALO_SEGS : 32;
        const long long g = ((long long)(t / PERIOD) * G + cta) * PB + j;    // global sub-tile (j == PB: head of the next run)
        uint32_t bytes = nseg * C::SEG_BYTES + (slot == 0 ? C::HALO_SEGS * C::SEG_BYTES : 0);
        uint32_t bar = bar_full + 8 * slot;
        if (lane == 0) mbar_expect_tx(bar, bytes);
        __syncwarp();
        const unsigned char *src = gin + g * SUB_BYTES + lane * C::SEG_BYTES;
        if (lane < nseg) bulk_g2s(ring + slot * C::SLOT_BYTES + lane * C::SEG_STRIDE, src, C::SEG_BYTES, bar);
        if (slot == 0 && lane < C::HALO_SEGS)
            bulk_g2s(ring + C::NS * C::SLOT_BYTES + lane * C::SEG_STRIDE, src, C::SEG_BYTES, bar);
        if (lane == 0) gen_publish(gen_armed + 4 * slot, t / C::NS + 1);
    };
    // Make sure tile t's fill gets issued, by exactly one warp: whoever moves the slot's claim word from gen-1 to gen.
    // The slots are shared between warps (13 slots, 8 warps), so the caller may be AHEAD of the slot's history:
    //   * the previous generation may not even have been claimed yet (its consumer is behind): nobody else is then
    //     guaranteed to come back for this tile, so the caller waits for that claim and takes this one itself;
    //   * the previous generation may be claimed but not landed (its claimer still waits for ITS predecessor's readers).
    //     The one-bit parity wait on the `empty` barrier below would then alias with the phase two back, return at once
    //     and let this fill overwrite a slot that is still being read -- and arrive a second time on a `full` barrier
    //     whose phase is still open (an over-arrival: the kernel dies with "unspecified launch failure").  So first wait,
    //     generation-guarded, until the previous generation HAS landed; from then on the `empty` barrier is in the phase
    //     the parity names.
    // In the steady state the claim was made long ago by the warp that freed the slot and all of this is one shared load.
    auto claim_and_fill = [&](int t) {
        const int slot = t % C::NS, gen = t / C::NS + 1;
        int won = 0;
        if (lane == 0) {
            for (;;) {
                const int c = *(volatile int *)&s_claim[slot];
                if (c >= gen) break;                                          // claimed (the claimer issues the fill)
                if (c == gen - 1) { if (atomicCAS(&s_claim[slot], gen - 1, gen) == gen - 1) { won = 1; break; } continue; }
                __nanosleep(64);                                              // the previous generation is still unclaimed
            }
        }
        won = __shfl_sync(0xffffffffu, won, 0);
        if (!won) return;
        if (gen > 1) {
            slot_wait<true>(bar_full + 8 * slot, gen_armed + 4 * slot, gen - 1);   // the previous generation has landed ...
            mbar_wait(bar_empty + 8 * slot, (gen - 2) & 1);                        // ... and has been consumed
        }
        issue_fill(t);
    };

    float tap[T];
#pragma unroll
    for (int k = 0; k < T; k++) tap[k] = __ldg(taps + k);

    // prologue: fill the ring with whatever is already published (never waits)
    for (int t = warp; t < C::NS; t += C::NWARPS)
        if (run_of(t) < runs_total && have_run(run_of(t), false)) claim_and_fill(t);

    // Run bookkeeping is deferred by one tile: the fence that orders a tile's output stores before the run counter is
    // cheap once the stores have drained (a tile later) and expensive right behind them.  Before a warp waits for input
    // it publishes what it holds, so a stalled stream still sees every finished run.
    int pend_q = -1;
    long long pend_r = 0;
    auto publish_pending = [&]() {
        if (pend_q < 0) return;
        if (lane == 0) {
            // CTA scope is enough here: the warps of a run synchronise through the shared-memory counter, and the one that
            // completes the run issues the (cumulative) system-scope fence before the flag
            __threadfence_block();
            if (atomicAdd(&s_run_cnt[pend_q & 7], 1) == C::NWARPS - 1) {
                s_run_cnt[pend_q & 7] = 0;
                __threadfence_system();
                ctl->done[pend_r] = 1u;
            }
        }
        pend_q = -1;
    };

    for (int t = warp; ; t += C::NWARPS) {
        const int q = t / PERIOD, j = t % PERIOD;
        const long long r = run_of(t);
        if (r >= runs_total) break;                                   // out of the session's capacity
        if (!have_run(r, false)) {
            publish_pending();
            if (!have_run(r, true)) break;                            // closed: this run will never exist
        }
        claim_and_fill(t);                                             // unless the warp that freed the slot already did
        const int slot = t % C::NS, slot2 = (t + 1) % C::NS;
        if (j == PB) {
            // halo-only slot: nothing to compute; the consumer of t-1 arrives for having read it, this warp for "owning" it.
            // The arrival must count for THIS generation: wait until its fill has landed first (another warp may have claimed
            // the fill and still be waiting for the previous generation's readers -- an early arrival would complete their
            // phase for them and let the slot be overwritten under a reader).
            slot_wait<true>(bar_full + 8 * slot, gen_armed + 4 * slot, t / C::NS + 1);
            if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
        } else {
            slot_wait<true>(bar_full + 8 * slot, gen_armed + 4 * slot, t / C::NS + 1);
            slot_wait<true>(bar_full + 8 * slot2, gen_armed + 4 * slot2, (t + 1) / C::NS + 1);
            const unsigned char *base = smem + slot * C::SLOT_BYTES + lane * C::SEG_STRIDE;
            u64 acc[R];
#pragma unroll
            for (int rr = 0; rr < R; rr++) acc[rr] = 0ULL;
#pragma unroll
            for (int c = 0; c < C::NCH; c++) {
                const int e0 = c * 2;
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(base + (e0 / C::SEG_ELEMS) * C::SEG_STRIDE + (e0 % C::SEG_ELEMS) * 8);
#pragma unroll
                for (int rr = 0; rr < R; rr++) {
                    const int k0 = e0 - rr * D, k1 = e0 + 1 - rr * D;
                    if (k0 >= 0 && k0 < T) acc[rr] = ffma2(v.x, dup2(tap[k0 < 0 ? 0 : (k0 >= T ? 0 : k0)]), acc[rr]);
                    if (k1 >= 0 && k1 < T) acc[rr] = ffma2(v.y, dup2(tap[k1 < 0 ? 0 : (k1 >= T ? 0 : k1)]), acc[rr]);
                }
            }
            const long long m0 = (r * PB + j) * (long long)C::SUB_OUT + lane * R;
            u64 *os = reinterpret_cast<u64 *>(out) + m0;
            if (vec_store) {
                ulonglong2 *o = reinterpret_cast<ulonglong2 *>(os);
#pragma unroll
                for (int rr = 0; rr < R; rr += 2) o[rr / 2] = make_ulonglong2(acc[rr], acc[rr + 1]);
            } else {
#pragma unroll
                for (int rr = 0; rr < R; rr++) os[rr] = acc[rr];
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_empty + 8 * slot);
                if (j == 0) mbar_arrive(bar_empty + 8 * slot);   // the first sub-tile of a run has no predecessor using it as halo
                mbar_arrive(bar_empty + 8 * slot2);
            }
            // A run's 32 sub-tiles fall to the 8 warps four each; a warp reports once per run, behind its LAST sub-tile of it:
            // one CTA-scope fence and one shared-memory atomic per warp and run instead of one per sub-tile (no measurable
            // difference in throughput -- the deferred fence was already cheap -- but a quarter of the bookkeeping).
            publish_pending();          // what this warp finished one tile ago
            static_assert(PB % C::NWARPS == 0, "every warp owns PB / NWARPS sub-tiles of a run");
            if (j >= PB - C::NWARPS) { pend_q = q; pend_r = r; }
        }
        // opportunistic refill of the slot just freed -- only if the target's run is already published (never waits for input)
        const int t2 = t + C::NS;
        const long long r2 = run_of(t2);
        if (!(dbg_flags & 1) && r2 < runs_total && have_run(r2, false)) claim_and_fill(t2);
    }
    publish_pending();
}

// geometry of the persistent consumer that serves a shape: samples per run and halo samples behind a run; false: none
bool dec_persist_geometry(int taps_stored, int D, bool cplx, int *run_samples, int *halo_samples) {
    if (!cplx || D != 8 || taps_stored > 128 || taps_stored < 1) return false;
    *run_samples = 32 * 256 * 8;
    *halo_samples = taps_stored <= 32 ? RingCfg<true, 32, 8, 8>::HALO_SEGS * 64
                  : taps_stored <= 64 ? RingCfg<true, 64, 8, 8>::HALO_SEGS * 64 : RingCfg<true, 128, 8, 8>::HALO_SEGS * 64;
    return true;
}

template <int T>
static int launch_persist_t(Ctx *c, const float *d_taps, const void *d_in, void *d_out, void *ctl, void *d_relay, long long runs_total,
                            cudaStream_t stream, int grid) {
    typedef RingCfg<true, T, 8, 8> C;
    SDR_TRY(ring_attr(c, reinterpret_cast<const void *>(k_dec_ring_persist<true, T, 8, 8, 32>), C::SMEM_BYTES));
    // test knob, bit 0: no opportunistic refills -- every fill is then issued by the warp that consumes the tile, the path
    // the kernel otherwise takes only when it runs ahead of the host (tests/test_gpu_persistent.py runs the suite that way too)
    static const int dbg_flags = getenv("SDR_B200_PERSIST_FLAGS") ? atoi(getenv("SDR_B200_PERSIST_FLAGS")) : 0;
    k_dec_ring_persist<true, T, 8, 8, 32><<<grid + 1, 256, C::SMEM_BYTES, stream>>>(d_in, d_out, d_taps, (PersistCtl *)ctl,
                                                                                     (PersistRelay *)d_relay, runs_total, dbg_flags);
    return SDR_OK;
}

// launches the persistent consumer on `stream` over the contiguous device stream at d_in: grid = SMs - 8 worker CTAs (the
// remaining SMs stay free for whatever else the process launches while the consumer is resident) + 1 relay CTA.  A run =
// 32 sub-tiles = 8192 outputs = one yielded vector of the headline configuration.
int launch_dec_persist(Ctx *c, int taps_stored, int D, bool cplx, const float *d_taps, const void *d_in, void *d_out, void *ctl,
                       void *d_relay, long long runs_total, cudaStream_t stream, int *grid_out, const char **name) {
    *name = "none";
    int rs, hs;
    if (!dec_persist_geometry(taps_stored, D, cplx, &rs, &hs)) return set_error(SDR_EINVAL, "no persistent kernel for this shape");
    SDR_TRY(c->bind());
    int grid = c->sm_count - 8;
    if (grid < 1) grid = 1;
    *grid_out = grid;
    SDR_CUDA(cudaMemsetAsync(d_relay, 0, sizeof(PersistRelay), stream));
    if (taps_stored <= 32)      { *name = "dec_c_ring_persist<32,8,8,32>";  SDR_TRY(launch_persist_t<32>(c, d_taps, d_in, d_out, ctl, d_relay, runs_total, stream, grid)); }
    else if (taps_stored <= 64) { *name = "dec_c_ring_persist<64,8,8,32>";  SDR_TRY(launch_persist_t<64>(c, d_taps, d_in, d_out, ctl, d_relay, runs_total, stream, grid)); }
    else                        { *name = "dec_c_ring_persist<128,8,8,32>"; SDR_TRY(launch_persist_t<128>(c, d_taps, d_in, d_out, ctl, d_relay, runs_total, stream, grid)); }
    c->launches++;
    SDR_CUDA(cudaGetLastError());
    return SDR_OK;
}

}  // namespace sdr
