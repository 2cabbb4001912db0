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
#include "ring_common.cuh"

#include <cstdlib>

namespace sdr {

// One instantiation per (data type, stored tap count T, decimation D, outputs per lane R).  T is the kernel's tap
// capacity: a record with fewer taps runs on the next larger instantiation with its tap array zero-padded (FirRec keeps
// d_taps padded to 128 floats), so e.g. the FM example's 51-tap RF decimator (examples/fm/Coeffs.hs:11-66) is the <64, 8>
// kernel.  T need not be a multiple of D.
template <bool CPLX, int T, int D, int R, int NW = 8>
struct RingCfg {
    static constexpr int EB = CPLX ? 8 : 4;                 // bytes per stream element
    static constexpr int EPC = 16 / EB;                     // elements per 16-byte chunk (one LDS.128)
    static constexpr int SEG_ELEMS = R * D;                 // the input elements a lane's R outputs advance over
    static constexpr int SEG_BYTES = SEG_ELEMS * EB;
    static_assert(SEG_BYTES % 16 == 0, "lane segments are moved by 16-byte bulk copies");
    static constexpr int SEG_STRIDE = SEG_BYTES + 16;       // +16 B: lanes' LDS.128 land on distinct bank groups
    static constexpr int SUB_OUT = 32 * R;                  // outputs per sub-tile (one warp pass)
    static constexpr int SLOT_BYTES = 32 * SEG_STRIDE;
    static constexpr int WIN = (R - 1) * D + T;             // elements a lane reads
    static constexpr int NCH = (WIN + EPC - 1) / EPC;       // ... as 16-byte chunks
    static constexpr int HALO_RAW = (NCH * EPC - SEG_ELEMS + SEG_ELEMS - 1) / SEG_ELEMS;
    static constexpr int HALO_SEGS = HALO_RAW < 1 ? 1 : HALO_RAW;   // segments of the NEXT sub-tile a pass reads
    static constexpr int NWARPS = NW;
    static constexpr int NS_FIT = (220 * 1024 - HALO_SEGS * SEG_STRIDE - 256) / SLOT_BYTES;
    static constexpr int NS = NS_FIT >= 2 * NWARPS ? 2 * NWARPS : NS_FIT;                  // ring slots
    static constexpr bool GUARD = (NS % NWARPS) != 0;   // see slot_wait in ring_common.cuh
    static constexpr int RING_BYTES = NS * SLOT_BYTES + HALO_SEGS * SEG_STRIDE;           // + mirror of slot 0's head
    static constexpr int SMEM_BYTES = RING_BYTES + 2 * NS * 8 + NS * 4 + 256;
    static_assert(NS >= NWARPS + 3, "ring too small for the warps plus prefetch");
    static_assert(HALO_SEGS <= 32, "halo wider than a sub-tile");
    static_assert(CPLX ? R % 2 == 0 : R % 4 == 0, "outputs per lane are stored in 16-byte groups");
};

// TP: the taps travel as launch parameters (constant bank) instead of living in registers -- what makes 256 taps fit.
template <int N> struct TapBlock { float t[N]; };

template <bool CPLX, int T, int D, int R, bool TP = false, int NW = 8>
__global__ void __launch_bounds__(32 * NW, 1)
k_dec_ring(const void *__restrict__ in, long long a_bytes, const void *__restrict__ in_b, long long total_bytes,
           void *__restrict__ out, long long num, const float *__restrict__ taps, long long n_sub,
           const __grid_constant__ TapBlock<TP ? T : 1> K) {
    typedef RingCfg<CPLX, T, D, R, NW> C;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool vec_store = (reinterpret_cast<uintptr_t>(out) & 15) == 0;

    // contiguous, balanced range of sub-tiles for this CTA
    long long q = n_sub / gridDim.x, rem = n_sub % gridDim.x;
    long long s0 = blockIdx.x * q + (blockIdx.x < rem ? blockIdx.x : rem);
    int cnt = (int)(q + (blockIdx.x < rem ? 1 : 0));   // local sub-tiles 0..cnt-1 are computed; cnt is halo-only
    if (cnt == 0) return;

    // Programmatic dependent launch: let the next kernel in the stream be scheduled as this one's CTAs drain (its launch
    // latency, CTA start-up and barrier set-up then overlap our tail) ...
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t ring = smem_u32(smem);
    const uint32_t bar_full = ring + C::RING_BYTES + ((128 - C::RING_BYTES % 128) % 128);
    const uint32_t bar_empty = bar_full + C::NS * 8;
    const uint32_t gen_armed = bar_empty + C::NS * 8;   // generation guard, see ring_common.cuh
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NS; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 2);
                                          asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(gen_armed + 4 * s), "r"(0) : "memory"); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    // ... and, launched that way ourselves, touch global memory only after everything before us in the stream has
    // completed and flushed (a no-op when the predecessor did not trigger early): stream-order semantics are unchanged
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const unsigned char *gin = reinterpret_cast<const unsigned char *>(in);
    // The stream is `in` (a_bytes bytes) followed by `in_b` (up to total_bytes; the right neighbour's chunk on a sharded
    // pass, or nothing); everything beyond total_bytes reads as zero.  A fill that lies wholly inside `in` -- all of them
    // except the last one or two of the last CTA -- takes the fast path: one bulk copy of a whole segment per lane.
    constexpr long long SUB_BYTES = 32LL * C::SEG_BYTES;
    const long long cta_bytes = a_bytes - s0 * SUB_BYTES;
    const int fast_full = (int)(cta_bytes <= 0 ? 0 : (cta_bytes / SUB_BYTES > cnt ? cnt : cta_bytes / SUB_BYTES));
    const bool halo_fast = cta_bytes >= cnt * SUB_BYTES + C::HALO_SEGS * C::SEG_BYTES;
    auto issue_fill_edge = [&](int u) {
        const int slot = u % C::NS;
        const int nseg = (u == cnt) ? C::HALO_SEGS : 32;
        const long long start = (s0 + u) * SUB_BYTES;
        const uint32_t bar = bar_full + 8 * slot;
        const long long ls = start + lane * C::SEG_BYTES;                       // stream offset of this lane's segment
        long long v = total_bytes - ls;
        const int valid = lane < nseg ? (int)(v < 0 ? 0 : (v > C::SEG_BYTES ? C::SEG_BYTES : v)) : 0;
        const bool mirror = slot == 0 && lane < C::HALO_SEGS;
        const uint32_t dst = ring + slot * C::SLOT_BYTES + lane * C::SEG_STRIDE;
        const uint32_t dst_m = ring + C::NS * C::SLOT_BYTES + lane * C::SEG_STRIDE;
        if (lane < nseg)
            for (int o = valid; o < C::SEG_BYTES; o += 16) {                    // what the stream does not hold reads as zero
                asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst + o), "r"(0) : "memory");
                if (mirror) asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst_m + o), "r"(0) : "memory");
            }
        __syncwarp();
        if (lane == 0) {
            long long t = total_bytes - start;
            long long tx = t < 0 ? 0 : (t > nseg * C::SEG_BYTES ? nseg * C::SEG_BYTES : t);
            if (slot == 0) tx += t < 0 ? 0 : (t > C::HALO_SEGS * C::SEG_BYTES ? C::HALO_SEGS * C::SEG_BYTES : t);
            mbar_expect_tx(bar, (uint32_t)tx);
        }
        __syncwarp();
        if (valid > 0) {
            long long va = a_bytes - ls;
            const int from_a = (int)(va < 0 ? 0 : (va > valid ? valid : va));
            const unsigned char *src_b = reinterpret_cast<const unsigned char *>(in_b) + (ls + from_a - a_bytes);
            if (from_a > 0) {
                bulk_g2s(dst, gin + ls, from_a, bar);
                if (mirror) bulk_g2s(dst_m, gin + ls, from_a, bar);
            }
            if (valid > from_a) {
                bulk_g2s(dst + from_a, src_b, valid - from_a, bar);
                if (mirror) bulk_g2s(dst_m + from_a, src_b, valid - from_a, bar);
            }
        }
        if (lane == 0) gen_publish(gen_armed + 4 * slot, u / C::NS + 1);
    };
    auto issue_fill = [&](int u) {
        if (!(u < fast_full || (u == cnt && halo_fast))) { issue_fill_edge(u); return; }
        int slot = u % C::NS;
        int nseg = (u == cnt) ? C::HALO_SEGS : 32;
        uint32_t bytes = nseg * C::SEG_BYTES + (slot == 0 ? C::HALO_SEGS * C::SEG_BYTES : 0);
        uint32_t bar = bar_full + 8 * slot;
        if (lane == 0) mbar_expect_tx(bar, bytes);
        __syncwarp();
        const unsigned char *src = gin + (s0 + u) * SUB_BYTES + lane * C::SEG_BYTES;
        if (lane < nseg) bulk_g2s(ring + slot * C::SLOT_BYTES + lane * C::SEG_STRIDE, src, C::SEG_BYTES, bar);
        if (slot == 0 && lane < C::HALO_SEGS)
            bulk_g2s(ring + C::NS * C::SLOT_BYTES + lane * C::SEG_STRIDE, src, C::SEG_BYTES, bar);
        if (lane == 0) gen_publish(gen_armed + 4 * slot, u / C::NS + 1);
    };

    for (int u = warp; u < C::NS && u <= cnt; u += C::NWARPS) issue_fill(u);

    float tap[TP ? 1 : T];
    if (!TP) {
#pragma unroll
        for (int k = 0; k < (TP ? 1 : T); k++) tap[k] = __ldg(taps + k);
    }
#define SDR_TAP(k) (TP ? K.t[TP ? (k) : 0] : tap[TP ? 0 : (k)])

    for (int u = warp; u < cnt; u += C::NWARPS) {
        const int slot = u % C::NS, par = (u / C::NS) & 1;
        const int slot2 = (u + 1) % C::NS;
        slot_wait<C::GUARD>(bar_full + 8 * slot, gen_armed + 4 * slot, u / C::NS + 1);
        slot_wait<C::GUARD>(bar_full + 8 * slot2, gen_armed + 4 * slot2, (u + 1) / C::NS + 1);

        const unsigned char *base = smem + slot * C::SLOT_BYTES + lane * C::SEG_STRIDE;
        const long long m0 = (s0 + u) * (long long)C::SUB_OUT + lane * R;
        if (CPLX) {
            u64 acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = 0ULL;
#pragma unroll
            for (int c = 0; c < C::NCH; c++) {
                const int e0 = c * 2;
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(base + (e0 / C::SEG_ELEMS) * C::SEG_STRIDE + (e0 % C::SEG_ELEMS) * 8);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int k0 = e0 - r * D, k1 = e0 + 1 - r * D;
                    if (k0 >= 0 && k0 < T) acc[r] = ffma2(v.x, dup2(SDR_TAP(k0 < 0 ? 0 : (k0 >= T ? 0 : k0))), acc[r]);
                    if (k1 >= 0 && k1 < T) acc[r] = ffma2(v.y, dup2(SDR_TAP(k1 < 0 ? 0 : (k1 >= T ? 0 : k1))), acc[r]);
                }
            }
            u64 *os = reinterpret_cast<u64 *>(out) + m0;
            if (m0 + R > num) {   // ragged last sub-tile of the stream
#pragma unroll
                for (int r = 0; r < R; r++) if (m0 + r < num) os[r] = acc[r];
            } else if (vec_store) {
                ulonglong2 *o = reinterpret_cast<ulonglong2 *>(os);
#pragma unroll
                for (int r = 0; r < R; r += 2) o[r / 2] = make_ulonglong2(acc[r], acc[r + 1]);
            } else {   // output only 8-byte aligned (a pipe's FIFO cursor after an odd number of outputs)
#pragma unroll
                for (int r = 0; r < R; r++) os[r] = acc[r];
            }
        } else {
            float acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = 0.0f;
#pragma unroll
            for (int c = 0; c < C::NCH; c++) {
                const int e0 = c * 4;
                const float4 v = *reinterpret_cast<const float4 *>(base + (e0 / C::SEG_ELEMS) * C::SEG_STRIDE + (e0 % C::SEG_ELEMS) * 4);
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int k = e0 + i - r * D;
                        if (k >= 0 && k < T) acc[r] = fmaf(SDR_TAP(k < 0 ? 0 : (k >= T ? 0 : k)), e[i], acc[r]);
                    }
                }
            }
            float *os = reinterpret_cast<float *>(out) + m0;
            if (m0 + R > num) {
#pragma unroll
                for (int r = 0; r < R; r++) if (m0 + r < num) os[r] = acc[r];
            } else if (vec_store) {
                float4 *o = reinterpret_cast<float4 *>(os);
#pragma unroll
                for (int r = 0; r < R; r += 4) o[r / 4] = make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) os[r] = acc[r];
            }
        }

        __syncwarp();
        if (lane == 0) {
            mbar_arrive(bar_empty + 8 * slot);
            if (u == 0) mbar_arrive(bar_empty + 8 * slot);   // sub-tile 0 has no predecessor using it as halo
            mbar_arrive(bar_empty + 8 * slot2);
        }
        if (u + C::NS <= cnt) {
            mbar_wait(bar_empty + 8 * slot, par);
            issue_fill(u + C::NS);
        }
    }
}

template <bool CPLX, int T, int D, int R, bool TP = false, int NW = 8>
static int launch_ring(Ctx *c, const float *d_taps, Seg2 seg, void *d_out, long long num, long long *done, const float *h_taps = nullptr) {
    typedef RingCfg<CPLX, T, D, R, NW> C;
    constexpr int EB = C::EB, EPC = C::EPC;
    const long long n_in = seg.na + seg.nb;
    const long long needed = (num - 1) * D + T;                    // elements the `num` outputs read (T = the kernel's tap capacity)
    const long long needed2 = (needed + EPC - 1) / EPC * EPC;      // bulk copies move whole 16-byte units
    const long long usable = n_in / EPC * EPC;
    long long n_sub, a_bytes, total_bytes;
    // COVERING mode: the kernel produces all `num` outputs, ragged last sub-tile and windows that run into the second
    // segment included (its edge fills split a lane segment between the two sources and zero-fill what lies beyond).
    // It needs both sources and the boundary between them on 16-byte boundaries.  The caller's outputs are valid for the
    // record's own tap count, which may be smaller than T: windows may then reach up to T - taps elements past the
    // resident data, where the zero-padded taps meet zero-filled shared memory.
    const bool covering = num > 0 && (seg.na % EPC) == 0 && (seg.nb == 0 || (((uintptr_t)seg.b) & 15) == 0) && (n_in % EPC) == 0;
    if (covering) {
        n_sub = (num + C::SUB_OUT - 1) / C::SUB_OUT;
        a_bytes = seg.na * EB;
        total_bytes = usable * EB;
        if (total_bytes < a_bytes) a_bytes = total_bytes;
        *done = num;
    } else {
        // interior only: the sub-tiles whose whole window (halo segments included) is resident in the first segment
        long long by_in = (seg.na / C::SEG_ELEMS - C::HALO_SEGS) / 32;
        n_sub = num / C::SUB_OUT;
        if (by_in < n_sub) n_sub = by_in;
        if (n_sub <= 0) { *done = 0; return SDR_OK; }
        a_bytes = total_bytes = (seg.na * EB) & ~15LL;
        *done = n_sub * C::SUB_OUT;
    }
    (void)needed2;
    SDR_TRY(c->bind());
    SDR_TRY(ring_attr(c, reinterpret_cast<const void *>(k_dec_ring<CPLX, T, D, R, TP, NW>), C::SMEM_BYTES));
    int sms = c->sm_count - c->reserve_sms;
    if (sms < 1) sms = 1;
    int grid = (int)(n_sub < sms ? n_sub : sms);
    const long long num_mask = covering ? num : n_sub * C::SUB_OUT;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32 * NW); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = c->s();
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        static const bool no_pdl = getenv("SDR_B200_NO_PDL") != nullptr;   // measurement knob
        cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
        const void *a0 = seg.a, *b0 = seg.b;
        TapBlock<TP ? T : 1> K = {};
        if (TP) for (int k = 0; k < T; k++) K.t[TP ? k : 0] = h_taps[k];
        SDR_CUDA(cudaLaunchKernelEx(&cfg, k_dec_ring<CPLX, T, D, R, TP, NW>, a0, a_bytes, b0, total_bytes, d_out, num_mask, d_taps, n_sub, K));
    }
    c->launches++;
    SDR_CUDA(cudaGetLastError());
    return SDR_OK;
}

// true when launch_dec_fast would produce ALL `num` outputs in one launch (no generic tail to fork): same conditions as
// the covering mode of launch_ring
bool dec_fast_will_cover(bool cplx, int taps_stored, int D, Seg2 seg, long long num) {
    if (num <= 0 || taps_stored > 256 || (D != 4 && D != 8 && D != 16)) return false;
    const int epc = cplx ? 2 : 4;
    return (((uintptr_t)seg.a) & 15) == 0 && (seg.na % epc) == 0 && (seg.nb == 0 || (((uintptr_t)seg.b) & 15) == 0) &&
           ((seg.na + seg.nb) % epc) == 0;
}

// x = seg.a ++ seg.b.  taps_stored = the record's tap count (d_taps is zero-padded to at least 128 floats).
int launch_dec_fast(Ctx *c, bool cplx, int taps_stored, int D, const float *d_taps, Seg2 seg, void *d_out, long long num,
                    long long *done, const char **name, const float *h_taps) {
    *done = 0;
    *name = cplx ? "fir_direct" : "fir_tile";
    if ((((uintptr_t)seg.a) & 15) != 0) return SDR_OK;   // TMA bulk copies need a 16-byte aligned source; any output alignment
    if (taps_stored > 128) {
        // 129..256 taps: the taps cannot live in registers -- they travel as launch parameters and reach every FFMA / FFMA2
        // as a uniform-register operand (h_taps: the record's host copy, zero-padded here).  78 registers instead of 168.
        // FP32-pipe bound at decimation 4 and 8 (128 / 64 FMA per complex input sample), HBM-bound at 16.
        if (!(taps_stored <= 256 && h_taps && (D == 4 || D == 8 || D == 16))) return SDR_OK;
        float padded[256] = {0.0f};
        for (int k = 0; k < taps_stored; k++) padded[k] = h_taps[k];
#define SDR_RING_P(CP, DD, RR, label)                                                     \
        if (cplx == CP && D == DD) { *name = label; return launch_ring<CP, 256, DD, RR, true>(c, d_taps, seg, d_out, num, done, padded); }
        SDR_RING_P(true, 8, 8, "dec_c_ring<256,8,8,param>")
        SDR_RING_P(true, 4, 8, "dec_c_ring<256,4,8,param>")
        SDR_RING_P(true, 16, 4, "dec_c_ring<256,16,4,param>")
        SDR_RING_P(false, 8, 8, "dec_r_ring<256,8,8,param>")
        SDR_RING_P(false, 4, 8, "dec_r_ring<256,4,8,param>")
        SDR_RING_P(false, 16, 8, "dec_r_ring<256,16,8,param>")
#undef SDR_RING_P
        return SDR_OK;
    }
    const int T = taps_stored <= 32 ? 32 : taps_stored <= 64 ? 64 : 128;
    // Real data, 128 taps, decimation 4 / 8: taps as launch parameters (uniform-register operands), which frees the registers
    // for 16 warps at 8 outputs per lane -- 588 / 1055 Gsamples/s against 506 / 969 for the register-tap form below (which
    // stays for 64 taps, where it is the faster one: 824 / 1486 against 760 / 1397).  SDR_B200_DEC_TP=0 switches it off,
    // =2 also takes the 64-tap shapes (measurement knob).
    static const int tp_mode = getenv("SDR_B200_DEC_TP") ? atoi(getenv("SDR_B200_DEC_TP")) : 1;
    if (tp_mode && h_taps && !cplx && (D == 4 || D == 8) && (T == 128 || (tp_mode == 2 && T == 64))) {
        float padded[128] = {0.0f};
        for (int k = 0; k < taps_stored; k++) padded[k] = h_taps[k];
        if (T == 128 && D == 8) { *name = "dec_r_ring<128,8,8,param,16w>"; return launch_ring<false, 128, 8, 8, true, 16>(c, d_taps, seg, d_out, num, done, padded); }
        if (T == 128 && D == 4) { *name = "dec_r_ring<128,4,8,param,16w>"; return launch_ring<false, 128, 4, 8, true, 16>(c, d_taps, seg, d_out, num, done, padded); }
        if (T == 64 && D == 4) { *name = "dec_r_ring<64,4,8,param,16w>"; return launch_ring<false, 64, 4, 8, true, 16>(c, d_taps, seg, d_out, num, done, padded); }
        if (T == 64 && D == 8) { *name = "dec_r_ring<64,8,8,param,16w>"; return launch_ring<false, 64, 8, 8, true, 16>(c, d_taps, seg, d_out, num, done, padded); }
    }
#define SDR_RING(CP, TT, DD, RR, label)                                                     \
    if (cplx == CP && T == TT && D == DD) { *name = label; return launch_ring<CP, TT, DD, RR>(c, d_taps, seg, d_out, num, done); }
    SDR_RING(true, 128, 8, 8, "dec_c_ring<128,8,8>")
    SDR_RING(true, 64, 8, 8, "dec_c_ring<64,8,8>")
    SDR_RING(true, 32, 8, 8, "dec_c_ring<32,8,8>")
    SDR_RING(true, 128, 4, 16, "dec_c_ring<128,4,16>")
    SDR_RING(true, 64, 4, 16, "dec_c_ring<64,4,16>")
    SDR_RING(true, 32, 4, 16, "dec_c_ring<32,4,16>")
    SDR_RING(true, 128, 16, 4, "dec_c_ring<128,16,4>")
    SDR_RING(true, 64, 16, 4, "dec_c_ring<64,16,4>")
    SDR_RING(true, 32, 16, 4, "dec_c_ring<32,16,4>")
    SDR_RING(false, 128, 8, 16, "dec_r_ring<128,8,16>")
    SDR_RING(false, 64, 8, 16, "dec_r_ring<64,8,16>")
    SDR_RING(false, 32, 8, 16, "dec_r_ring<32,8,16>")
    SDR_RING(false, 128, 4, 32, "dec_r_ring<128,4,32>")
    SDR_RING(false, 64, 4, 32, "dec_r_ring<64,4,32>")
    SDR_RING(false, 32, 4, 32, "dec_r_ring<32,4,32>")
    SDR_RING(false, 128, 16, 8, "dec_r_ring<128,16,8>")
    SDR_RING(false, 64, 16, 8, "dec_r_ring<64,16,8>")
    SDR_RING(false, 32, 16, 8, "dec_r_ring<32,16,8>")
#undef SDR_RING
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
    __shared__ int s_run_cnt[8];         // sub-tiles stored per run in flight (indexed by q % 8)
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
            if (atomicAdd(&s_run_cnt[pend_q & 7], 1) == PB - 1) {
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
            publish_pending();          // the PREVIOUS tile of this warp
            pend_q = q; pend_r = r;
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
