// kernels_fm.cu -- fused front end of the FM receiver chain (cfg4, reference examples/fm/fm.hs:34-37):
//
//     P.map interleavedIQUnsignedByteToFloat >-> firDecimator deci >-> fmDemod
//
// as ONE persistent kernel: u8 IQ bytes in (2 B per sample instead of the 8 B of a float stream), float phase out
// (4 B per D samples), nothing in between touches HBM.  Structure and inner product are those of k_dec_c_ring
// (kernels_fast.cu); what changes:
//   * the ring holds raw bytes: a sub-tile (256 outputs = 2048 samples) is 4 KB, staged by coalesced 16-byte cp.async
//     (LDGSTS) into lane segments of 128 B padded to 144 B (odd number of 16-byte chunks: conflict-free LDS.128);
//     completion through cp.async.mbarrier.arrive on the slot's barrier.  24 slots = 3 per warp, so every generation
//     of a slot is consumed by the same warp and the parity wait needs no generation guard.
//   * bytes become floats in registers: PRMT drops the byte into the mantissa of 2^23 (0x4B0000bb = 8388608 + b), one
//     FFMA2 then computes (8388608 + b) * (1/128) - 65537 = (b - 128) / 128 for I and Q together -- exact, i.e.
//     bit-identical to convertC (reference convert.c:15-20).
//   * epilogue: phase(y[m] * conj(y[m-1])) per output (demod.cuh); y[m-1] of a lane's first output comes from the
//     previous lane by shuffle; the first output of a sub-tile needs the last output of the previous sub-tile, which
//     another warp computes at another time: every sub-tile therefore also stores its first and last COMPLEX output
//     (16 B per 2048 samples) and the 1-in-256 outputs are patched at the end of the kernel -- the boundaries inside a
//     CTA's own range by that CTA, the 147 boundaries between CTAs by whichever CTA takes the last ticket (fence +
//     atomic counter: by then every other CTA's boundary samples are visible).  Round 1 used a second launch for this:
//     3.4 us plus a launch gap behind every 39 us push.
// The FIR sum order is the same as everywhere else, so fused == un-fused bit for bit (tests/test_gpu_parity.py).
// Roofline: 2 B in + 0.5 B out per sample make HBM irrelevant; the kernel is FP32-pipe bound (1024 + 184 FFMA2 per pass).
#include "demod.cuh"
#include "ring_common.cuh"

#include <cstdio>
#include <cstdlib>

namespace sdr {

// tools/fm_timeline.cu compiles this file with SDR_FM_TIMING to get a per-warp time line of one launch (measurement aid;
// the library is built without it)
#ifdef SDR_FM_TIMING
__device__ unsigned long long g_fm_timing[160 * 16 * 8];
#define FM_T(i) do { if ((threadIdx.x & 31) == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
                     g_fm_timing[(blockIdx.x * 16 + (threadIdx.x >> 5)) * 8 + (i)] = t_; } } while (0)
#else
#define FM_T(i) do { } while (0)
#endif

template <int T, int D, int R, int NW>
struct FmCfg {
    static_assert(T % D == 0 && D == 8 && R == 8, "one decimation block = 8 IQ pairs = one 16-byte chunk");
    static constexpr int JB = T / D;
    static constexpr int BLK_BYTES = D * 2;                 // 16
    static constexpr int SEG_BYTES = R * BLK_BYTES;         // 128
    static constexpr int SEG_STRIDE = SEG_BYTES + 16;       // 144 = 9 chunks (odd)
    static constexpr int SUB_OUT = 32 * R;                  // 256 outputs
    static constexpr int SUB_BYTES = 32 * SEG_BYTES;        // 4096 input bytes
    static constexpr int SLOT_BYTES = 32 * SEG_STRIDE;      // 4608
    static constexpr int HALO_SEGS = (JB - 1 + R - 1) / R;  // 2
    static constexpr int NWARPS = NW;
    static constexpr int NS = 3 * NW <= 32 ? 3 * NW : 2 * NW;   // 24 slots for 8 warps, 32 for 16
    static constexpr int RING_BYTES = NS * SLOT_BYTES + HALO_SEGS * SEG_STRIDE;
    static constexpr int BAR_OFFSET = ((RING_BYTES + 127) / 128) * 128;
    static constexpr int SMEM_BYTES = BAR_OFFSET + 2 * NS * 8 + 128;
    static_assert(NS % NWARPS == 0, "same-warp slot ownership (no generation guard)");
};

// 16-byte asynchronous copy; src_bytes = 0 zero-fills the destination without touching global memory
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// byte `idx` (0..3) of `word` -> float 8388608 + b
__device__ __forceinline__ float byte_as_big_float(uint32_t word, int idx) {
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u | (uint32_t)idx));
}
__device__ __forceinline__ u64 pack2f(float lo, float hi) {
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ float2 unpack2f(u64 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

// bnd: 2 complex per sub-tile: [2t] = first output of sub-tile t, [2t+1] = last output.  The kernel covers ALL `num`
// outputs: chunks past the end of the stream are zero-filled (no valid output needs them), stores of the ragged last
// sub-tile are masked, and the stream's final complex output goes to *carry_out (the next call's carried sample).
// SYM: the taps are symmetric (c[k] = c[T-1-k], the usual linear-phase design): only T/2 of them are kept in registers,
// which halves the register footprint and lets 16 warps (instead of 8) share the SM -- the epilogue is latency bound,
// so the extra warps are what fills the FP32 pipe.  Same tap values in the same order: bit-identical results.
// DEMOD = false: the same kernel without the discriminator -- `P.map interleavedIQUnsignedByteToFloat >-> firDecimator`
// (fm.hs:34-36) alone: `out` then receives the decimated COMPLEX samples (8 B each), bnd / carry_out are unused.
template <int T, int D, int R, int NW, bool SYM, bool DEMOD = true>
__global__ void __launch_bounds__(32 * NW, 1)
k_fm_front_ring(const uint8_t *__restrict__ in, long long a_chunks, const uint8_t *__restrict__ in_b, long long n_chunks,
                float *__restrict__ out, long long num,
                float2 *__restrict__ bnd, float2 *__restrict__ carry_out, const float *__restrict__ taps, long long n_sub,
                const float2 *__restrict__ carry_in, unsigned int *__restrict__ ticket) {
    typedef FmCfg<T, D, R, NW> C;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool vec_store = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    long long q = n_sub / gridDim.x, rem = n_sub % gridDim.x;
    long long s0 = blockIdx.x * q + (blockIdx.x < rem ? blockIdx.x : rem);
    int cnt = (int)(q + (blockIdx.x < rem ? 1 : 0));   // >= 1: the grid never exceeds the number of sub-tiles

    FM_T(0);
    // programmatic dependent launch, as in k_dec_ring: the next launch in the stream may be scheduled while this one drains
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t ring = smem_u32(smem);
    const uint32_t bar_full = ring + C::BAR_OFFSET;
    const uint32_t bar_empty = bar_full + C::NS * 8;
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NS; s++) { mbar_init(bar_full + 8 * s, 32); mbar_init(bar_empty + 8 * s, 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // chunk c (16 B = 8 IQ pairs) of a sub-tile lands in segment c / 8 at offset (c % 8) * 16.  The byte stream is `in`
    // (a_chunks chunks: the stage's carried tail) followed by `in_b` (up to n_chunks in all: vectors read in place from
    // the caller's memory); chunks past n_chunks are zero-filled.
    auto chunk_src = [&](long long g) -> const unsigned char * {
        return g < a_chunks ? in + g * 16 : in_b + (g - a_chunks) * 16;
    };
    const uint32_t lane_dst = (uint32_t)((lane >> 3) * C::SEG_STRIDE + (lane & 7) * 16);   // chunk `lane` of a group of 32
    auto issue_fill = [&](int u) {
        const int slot = u % C::NS;
        const int nchunk = (u == cnt) ? C::HALO_SEGS * R : 32 * R;   // halo-only fill: the first segments
        const long long chunk0 = (s0 + u) * (long long)(32 * R);
        const uint32_t dst = ring + slot * C::SLOT_BYTES;
        // fast path (all fills but the one or two at a segment boundary / the end of the stream, and the halo-only one):
        // the whole sub-tile lies inside one source -- eight copies at constant offsets from two lane registers
        const bool in_a = chunk0 + 32 * R <= a_chunks, in_b_ = chunk0 >= a_chunks && chunk0 + 32 * R <= n_chunks;
        if (nchunk == 32 * R && (in_a || in_b_)) {
            const unsigned char *src = (in_a ? in + chunk0 * 16 : in_b + (chunk0 - a_chunks) * 16) + lane * 16;
#pragma unroll
            for (int i = 0; i < R; i++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + lane_dst + i * 4 * C::SEG_STRIDE), "l"(src + i * 512) : "memory");
            if (slot == 0 && lane < C::HALO_SEGS * R)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + C::NS * C::SLOT_BYTES + lane_dst), "l"(src) : "memory");
            cp_async_arrive(bar_full + 8 * slot);
            return;
        }
#pragma unroll
        for (int i = 0; i < R; i++) {
            const int c = i * 32 + lane;
            if (c < nchunk) {
                const bool ok = chunk0 + c < n_chunks;
                cp_async16(dst + (c >> 3) * C::SEG_STRIDE + (c & 7) * 16, ok ? chunk_src(chunk0 + c) : in, ok ? 16u : 0u);
            }
        }
        if (slot == 0 && lane < C::HALO_SEGS * R) {   // mirror of slot 0's head behind the last slot
            const bool ok = chunk0 + lane < n_chunks;
            cp_async16(ring + C::NS * C::SLOT_BYTES + (lane >> 3) * C::SEG_STRIDE + (lane & 7) * 16, ok ? chunk_src(chunk0 + lane) : in,
                       ok ? 16u : 0u);
        }
        cp_async_arrive(bar_full + 8 * slot);
    };

    constexpr int NT = SYM ? T / 2 : T;
    float tap[NT];
#pragma unroll
    for (int k = 0; k < NT; k++) tap[k] = __ldg(taps + k);   // written once when the record was made: safe before the wait
    // everything before us in the stream has completed and flushed from here on (no-op without an early trigger)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    FM_T(1);

    // One fill per warp up front; the second generation of slots (u + NW) is requested from inside the first iteration,
    // once the first data has landed: asked for at once, the whole 148 KB ring queues behind the memory system and every
    // warp sits in its second fill for 2 us before it computes anything (tools/fm_timeline.cu).
    static_assert(C::NS == 2 * C::NWARPS || C::NS == 3 * C::NWARPS, "slots per warp");
    if (warp <= cnt) issue_fill(warp);
    FM_T(2);
    const u64 k_scale = dup2(0.0078125f), k_bias = dup2(-65537.0f);

    for (int u = warp; u < cnt; u += C::NWARPS) {
        const int slot = u % C::NS, slot2 = (u + 1) % C::NS;
        mbar_wait(bar_full + 8 * slot, (u / C::NS) & 1);
        mbar_wait(bar_full + 8 * slot2, ((u + 1) / C::NS) & 1);
        if (u == warp) {   // first iteration: the rest of this warp's slots (never used before: no empty-wait)
            FM_T(3);
            for (int v = u + C::NWARPS; v < C::NS && v <= cnt; v += C::NWARPS) issue_fill(v);
        }

        const unsigned char *base = smem + slot * C::SLOT_BYTES + lane * C::SEG_STRIDE;
        u64 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0ULL;
#pragma unroll
        for (int b = 0; b < R + C::JB - 1; b++) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(base + (b / R) * C::SEG_STRIDE + (b % R) * C::BLK_BYTES);
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
            u64 v[D];
#pragma unroll
            for (int p = 0; p < D; p++) {   // sample p: bytes (2p, 2p+1) of the chunk = (I, Q)
                const uint32_t word = w[p >> 1];
                const int b0 = (p & 1) * 2;
                v[p] = ffma2(pack2f(byte_as_big_float(word, b0), byte_as_big_float(word, b0 + 1)), k_scale, k_bias);
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int j = b - r;
                if (j < 0 || j >= C::JB) continue;
#pragma unroll
                for (int p = 0; p < D; p++) {
                    const int k = j * D + p;
                    acc[r] = ffma2(v[p], dup2(tap[(SYM && k >= T / 2) ? T - 1 - k : k]), acc[r]);
                }
            }
        }
        // the slot's bytes are in registers: hand it back before the (long) epilogue
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(bar_empty + 8 * slot);
            if (u == 0) mbar_arrive(bar_empty + 8 * slot);
            mbar_arrive(bar_empty + 8 * slot2);
        }
        if (u + C::NS <= cnt) {
            mbar_wait(bar_empty + 8 * slot, (u / C::NS) & 1);
            issue_fill(u + C::NS);
        }

        const long long m0 = (s0 + u) * (long long)C::SUB_OUT + lane * R;   // this lane's first output index
        if (!DEMOD) {
            u64 *os = reinterpret_cast<u64 *>(out) + m0;
            if (vec_store && m0 + R <= num) {
                ulonglong2 *o = reinterpret_cast<ulonglong2 *>(os);
#pragma unroll
                for (int r = 0; r < R; r += 2) o[r / 2] = make_ulonglong2(acc[r], acc[r + 1]);
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) if (m0 + r < num) os[r] = acc[r];
            }
            continue;
        }
        // epilogue: FM discriminator
        float2 y[R];
#pragma unroll
        for (int r = 0; r < R; r++) y[r] = unpack2f(acc[r]);
        float2 prev;
        prev.x = __shfl_up_sync(0xffffffffu, y[R - 1].x, 1);
        prev.y = __shfl_up_sync(0xffffffffu, y[R - 1].y, 1);
        float ph[R];
        ph[0] = fm_phase(y[0], prev);   // lane 0's value is meaningless here: patched by the boundary pass at the end of the kernel
#pragma unroll
        for (int r = 1; r < R; r++) ph[r] = fm_phase(y[r], y[r - 1]);
        float *os = out + m0;
        if (vec_store && m0 + R <= num) {
            float4 *o = reinterpret_cast<float4 *>(os);
#pragma unroll
            for (int r = 0; r < R; r += 4) o[r / 4] = make_float4(ph[r], ph[r + 1], ph[r + 2], ph[r + 3]);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) if (m0 + r < num) os[r] = ph[r];
        }
        if (lane == 0) bnd[2 * (s0 + u)] = y[0];
        if (lane == 31) bnd[2 * (s0 + u) + 1] = y[R - 1];
#pragma unroll
        for (int r = 0; r < R; r++) if (m0 + r == num - 1) *carry_out = y[r];
    }
    FM_T(4);
    if (!DEMOD) return;

    // ---- the first output of every sub-tile: phase(first[t] * conj(last[t-1])), sub-tile 0 from the carried sample --------
    __shared__ unsigned int is_last;
    __threadfence();     // this thread's boundary samples and outputs, device-wide, before the ticket below
    __syncthreads();
    for (int u = threadIdx.x; u < cnt; u += blockDim.x) {
        const long long t = s0 + u;
        if (u == 0 && t != 0) continue;   // needs the previous CTA's last output
        const float2 prev = (t == 0) ? *carry_in : __ldcg(bnd + 2 * (t - 1) + 1);
        out[t * C::SUB_OUT] = fm_phase(__ldcg(bnd + 2 * t), prev);
    }
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (is_last) {       // every other CTA fenced its boundary samples before taking its ticket
        __threadfence();
        for (int b = 1 + (int)threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
            const long long t = b * q + (b < rem ? b : rem);   // first sub-tile of CTA b
            out[t * C::SUB_OUT] = fm_phase(__ldcg(bnd + 2 * t), __ldcg(bnd + 2 * (t - 1) + 1));
        }
        if (threadIdx.x == 0) *ticket = 0;   // for the next launch (stream order)
    }
    FM_T(5);
}

// One launch of the fused kernel for tap capacity TK (the record's taps zero-padded up to it), D = 8.
template <int TK, int NW, bool SYM, bool DEMOD>
static int launch_front_inst(Ctx *c, const float *d_taps, const uint8_t *d_in, long long n_samples, const uint8_t *d_in_b,
                             long long a_samples, float *d_out, long long num, float2 *d_bnd, float2 *d_carry_out, long long n_sub,
                             const float2 *d_carry_in, unsigned int *d_ticket) {
    typedef FmCfg<TK, 8, 8, NW> C;
    SDR_TRY(ring_attr(c, reinterpret_cast<const void *>(k_fm_front_ring<TK, 8, 8, NW, SYM, DEMOD>), C::SMEM_BYTES));
    int grid = (int)(n_sub < c->sm_count ? n_sub : c->sm_count);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32 * NW); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = c->s();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("SDR_B200_NO_PDL") != nullptr;   // measurement knob
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    const long long a_chunks = a_samples / 8, n_chunks = n_samples / 8;
    SDR_CUDA(cudaLaunchKernelEx(&cfg, k_fm_front_ring<TK, 8, 8, NW, SYM, DEMOD>, d_in, a_chunks, d_in_b, n_chunks, d_out, num, d_bnd,
                                d_carry_out, d_taps, n_sub, d_carry_in, d_ticket));
    c->launches++;
    SDR_CUDA(cudaGetLastError());
    return SDR_OK;
}

// picks the instantiation: 128 taps -- symmetric taps keep half of them in registers and run 16 warps, arbitrary taps 8
// warps; 64 / 32 tap capacity (e.g. the FM example's 51-tap RF decimator, examples/fm/Coeffs.hs:11-66, zero-padded) --
// 16 warps.  `label` receives the kernel's name without its prefix.
template <bool DEMOD>
static int launch_front_any(Ctx *c, int taps_stored, const float *d_taps, bool symmetric, const uint8_t *d_in, long long n_samples,
                            const uint8_t *d_in_b, long long a_samples, float *d_out, long long num, float2 *d_bnd, float2 *d_carry_out,
                            long long n_sub, const char **label, const float2 *d_carry_in = nullptr, unsigned int *d_ticket = nullptr) {
    if (taps_stored > 64) {
        if (symmetric && taps_stored == 128) { *label = "<128,8,8,sym,16w>"; return launch_front_inst<128, 16, true, DEMOD>(c, d_taps, d_in, n_samples, d_in_b, a_samples, d_out, num, d_bnd, d_carry_out, n_sub, d_carry_in, d_ticket); }
        *label = "<128,8,8>";
        return launch_front_inst<128, 8, false, DEMOD>(c, d_taps, d_in, n_samples, d_in_b, a_samples, d_out, num, d_bnd, d_carry_out, n_sub, d_carry_in, d_ticket);
    }
    if (taps_stored > 32) { *label = "<64,8,8,16w>"; return launch_front_inst<64, 16, false, DEMOD>(c, d_taps, d_in, n_samples, d_in_b, a_samples, d_out, num, d_bnd, d_carry_out, n_sub, d_carry_in, d_ticket); }
    *label = "<32,8,8,16w>";
    return launch_front_inst<32, 16, false, DEMOD>(c, d_taps, d_in, n_samples, d_in_b, a_samples, d_out, num, d_bnd, d_carry_out, n_sub, d_carry_in, d_ticket);
}

// Fused convert + decimate + demod of outputs [0, num) of a byte stream holding n_samples IQ pairs.  d_carry: previous
// stream sample (re, im) on the device; d_carry_out receives the last decimated complex output (a different word: the
// kernel reads the one while it writes the other); d_bnd: scratch of 2 complex per sub-tile (ceil(num / 256) sub-tiles);
// d_ticket: one zero-initialised word the launches of a stream share (the kernel leaves it zero).  *done = num when the
// shape has a tuned kernel.  T = the record's stored tap count (d_taps zero-padded to >= 128 floats).
int launch_fm_front(Ctx *c, int T, int D, const float *d_taps, bool symmetric, const uint8_t *d_in, long long a_samples,
                    const uint8_t *d_in_b, long long n_samples, float *d_out, long long num, float2 *d_bnd,
                    long long bnd_capacity_subtiles, const float2 *d_carry, float2 *d_carry_out, unsigned int *d_ticket,
                    long long *done, const char **name) {
    *done = 0;
    *name = "unfused";
    if (T > 128 || D != 8 || num <= 0) return SDR_OK;
    if ((((uintptr_t)d_in) & 15) != 0 || (((uintptr_t)d_out) & 3) != 0) return SDR_OK;
    if (a_samples < n_samples && ((a_samples & 7) != 0 || (((uintptr_t)d_in_b) & 15) != 0)) return SDR_OK;   // segment boundary on a chunk
    if ((num - 1) * D + T > n_samples) return set_error(SDR_EINVAL, "launch_fm_front: %lld outputs need more than %lld samples", num, n_samples);
    const long long n_sub = (num + 255) / 256;
    if (bnd_capacity_subtiles < n_sub) return set_error(SDR_EINVAL, "launch_fm_front: boundary scratch too small");
    SDR_TRY(c->bind());
    const char *label = "";
    SDR_TRY(launch_front_any<true>(c, T, d_taps, symmetric, d_in, n_samples, d_in_b, a_samples, d_out, num, d_bnd, d_carry_out, n_sub, &label,
                                   d_carry, d_ticket));
    static thread_local char nm[64];
    snprintf(nm, sizeof(nm), "fm_front_ring%s", label);
    *name = nm;
    *done = num;
    return SDR_OK;
}

// Fused convert + decimate (no demodulation) of outputs [0, num) of a byte stream holding n_samples IQ pairs: complex
// outputs.  *done = num when the shape has a tuned kernel, else 0 (the caller runs the two stages one after the other).
int launch_dec_u8(Ctx *c, int T, int D, const float *d_taps, bool symmetric, const uint8_t *d_in, long long a_samples,
                  const uint8_t *d_in_b, long long n_samples, float *d_out, long long num, long long *done, const char **name) {
    *done = 0;
    *name = "unfused";
    if (T > 128 || D != 8 || num <= 0) return SDR_OK;
    if ((((uintptr_t)d_in) & 15) != 0 || (((uintptr_t)d_out) & 7) != 0) return SDR_OK;
    if (a_samples < n_samples && ((a_samples & 7) != 0 || (((uintptr_t)d_in_b) & 15) != 0)) return SDR_OK;
    if ((num - 1) * D + T > n_samples) return set_error(SDR_EINVAL, "launch_dec_u8: %lld outputs need more than %lld samples", num, n_samples);
    const long long n_sub = (num + 255) / 256;
    SDR_TRY(c->bind());
    const char *label = "";
    SDR_TRY(launch_front_any<false>(c, T, d_taps, symmetric, d_in, n_samples, d_in_b, a_samples, d_out, num, nullptr, nullptr, n_sub, &label));
    static thread_local char nm[64];
    snprintf(nm, sizeof(nm), "dec_u8_ring%s", label);
    *name = nm;
    *done = num;
    return SDR_OK;
}

}  // namespace sdr
