// dec_ring.cuh -- tuned sm_100a kernels for the headline shapes.
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
// (The template lives in this header so that its instantiations can be spread over several translation units --
// kernels_fast_c.cu, _r.cu, _p.cu -- which compile in parallel; one unit took eleven minutes.)
#pragma once
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
__device__ __forceinline__ u64 dup2_bits(float v) { const unsigned int u = __float_as_uint(v); return ((u64)u << 32) | (u64)u; }

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
#define SDR_DUP(v) (TP ? dup2_bits(v) : dup2(v))

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
                    if (k0 >= 0 && k0 < T) acc[r] = ffma2(v.x, SDR_DUP(SDR_TAP(k0 < 0 ? 0 : (k0 >= T ? 0 : k0))), acc[r]);
                    if (k1 >= 0 && k1 < T) acc[r] = ffma2(v.y, SDR_DUP(SDR_TAP(k1 < 0 ? 0 : (k1 >= T ? 0 : k1))), acc[r]);
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
    static const int grid_override = getenv("SDR_B200_GRID") ? atoi(getenv("SDR_B200_GRID")) : 0;   // measurement knob
    if (grid_override > 0 && grid_override < grid) grid = grid_override;
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

// the instantiation groups (one translation unit each); *name receives a static string naming the instantiation
int launch_ring_complex(Ctx *c, int T, int D, const float *d_taps, Seg2 seg, void *d_out, long long num, long long *done, const char **name);
int launch_ring_real(Ctx *c, int T, int D, const float *d_taps, Seg2 seg, void *d_out, long long num, long long *done, const char **name);
// launch-parameter forms: T = 256 (both data types, decimation 4 / 8 / 16) and the 16-warp real decimators (T = 128 / 64, decimation 4 / 8);
// h_taps holds T floats (zero padded).  *done stays 0 when there is no such instantiation.
int launch_ring_param(Ctx *c, bool cplx, int T, int D, const float *d_taps, const float *h_taps, Seg2 seg, void *d_out, long long num,
                      long long *done, const char **name);

}  // namespace sdr
