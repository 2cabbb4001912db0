// kernels_real.cu -- tuned sm_100a kernels for REAL streams: the stride-1 FIR filter (K1 / cfg1: fastFilterSymR /
// fastFilterR -> filterAVXSymmetricRR / filterAVXRR, reference filter.c:37-68) and the rational polyphase resampler
// (K3 / cfg3: fastResamplerR -> resampleAVXRR, reference resample.c:70-87, tables FilterInternal.hs:297-319).
//
// Same persistent structure as the complex decimator (kernels_fast.cu): one CTA per SM, 8 warps, the input stream is
// staged once into a shared-memory ring by TMA bulk copies (one copy per contiguous slot here), completion on
// mbarriers, slots refilled by the warp that consumed them.  Per pass a lane owns R consecutive outputs (filter) or C
// whole L-output / M-input cycles (resampler); its window is read with LDS.128 at a lane stride that is an ODD number
// of 16-byte chunks, so the 8 lanes of a phase hit 8 distinct bank groups without padding.  Taps live in registers and
// every multiply-add is a fully unrolled FFMA with compile-time tap and accumulator indices, summed in increasing
// tap order -- the same order as the generic kernels, so results are bit-identical to them.
//
// Rooflines: the 64-tap filter needs 64 FMA per 8 algorithmic bytes -- it is FP32-pipe bound
// (about 0.64 of the HBM roofline at the measured FMA peak); the 3/10 resampler needs 9 FMA per 5.2 bytes -- HBM bound.
#include "fir_ffa.cuh"
#include "ring_common.cuh"

#include <cstdlib>

namespace sdr {

// ---------------------------------------------------------------------------------------------------------------
// FIR with a small stride (D = 1: the filters; D = 2), real or complex data, contiguous-slot ring
// ---------------------------------------------------------------------------------------------------------------
// A lane owns R consecutive outputs = R * D input elements; R is chosen so that the lane stride R * D * EB is an ODD
// number of 16-byte chunks (80 B for every shape instantiated below), which makes the lanes' LDS.128 conflict-free
// without padding and lets ONE bulk copy fill a whole contiguous slot.  (Larger strides cannot be made odd that way;
// they run on the padded-segment ring of kernels_fast.cu.)  Complex data: one FFMA2 per (sample, tap) as in the
// decimator -- the reference's duplicated-coefficient trick (Filter.hs:206) is not needed.
template <bool CPLX, int T, int D, int R, int S>
struct FirRCfg {
    static constexpr int EB = CPLX ? 8 : 4;
    static constexpr int EPC = 16 / EB;
    static constexpr int LANE_ELEMS = R * D;
    static constexpr int LANE_BYTES = LANE_ELEMS * EB;
    static_assert(LANE_BYTES % 16 == 0 && (LANE_BYTES / 16) % 2 == 1, "lane stride must be an odd number of 16-byte chunks");
    static constexpr int PASS_OUT = 32 * R;
    static constexpr int SLOT_OUT = PASS_OUT * S;            // outputs per slot
    static constexpr int SLOT_ELEMS = SLOT_OUT * D;          // input elements per slot
    static constexpr int SLOT_BYTES = SLOT_ELEMS * EB;
    static constexpr int WIN = (R - 1) * D + T;              // elements a lane pass reads
    static constexpr int NCH = (WIN + EPC - 1) / EPC;        // LDS.128 per lane pass
    static constexpr int HALO = NCH * EPC - LANE_ELEMS;      // elements a pass reads beyond its own lane segment
    static_assert(HALO > 0, "tap count must exceed the lane advance");
    static constexpr int HALO_BYTES = HALO * EB;
    static constexpr int NS = (220 * 1024 - HALO_BYTES - 512) / SLOT_BYTES;
    typedef ContigRing<SLOT_BYTES, HALO_BYTES, NS> Ring;
    static_assert(NS >= 11, "ring too small for 8 warps plus prefetch");
};

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

// PAIR (real data, D = 1): two ADJACENT OUTPUTS share one FFMA2 -- (y[2p], y[2p+1]) += c[k] * (x[2p+k], x[2p+k+1]) -- so every
// output still sums its taps in increasing order (bit-identical to the scalar form and to the generic kernel) while the
// FP32 instructions halve: the scalar kernel sits at the scalar-FFMA issue peak (59.7 TFLOP/s measured, profiles/r01_mb_fma.txt),
// FFMA2 reaches 67.  Sample pairs that start at an even index come straight out of the LDS.128 registers, the odd ones cost
// a register move each.
template <bool CPLX, int T, int D, int R, int S, bool PAIR = false>
__global__ void __launch_bounds__(256, 1)
k_fir_ring(const void *__restrict__ in_v, void *__restrict__ out_v, const float *__restrict__ taps, long long n_slots) {
    typedef FirRCfg<CPLX, T, D, R, S> C;
    static_assert(!PAIR || (!CPLX && D == 1 && R % 4 == 0), "output pairing is the real stride-1 form");
    // widest store a lane's R outputs allow (compile time) and the output pointer permits (run time): ONE vector variant
    // plus the scalar one per instantiation, so the compiler clones the unrolled loop body only twice
    constexpr bool CAN16 = (R * C::EB) % 16 == 0;
    constexpr bool CAN8 = !CAN16 && (R * C::EB) % 8 == 0;
    const bool vec_store = (reinterpret_cast<uintptr_t>(out_v) & (CAN16 ? 15 : 7)) == 0 && (CAN16 || CAN8);
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long q = n_slots / gridDim.x, rem = n_slots % gridDim.x;
    long long s0 = blockIdx.x * q + (blockIdx.x < rem ? blockIdx.x : rem);
    int cnt = (int)(q + (blockIdx.x < rem ? 1 : 0));
    if (cnt == 0) return;

    typename C::Ring ring;
    ring.init(smem, reinterpret_cast<const unsigned char *>(in_v) + s0 * (long long)C::SLOT_BYTES, cnt);
    ring.prologue(warp, lane, 8);

    float tap[T];
#pragma unroll
    for (int k = 0; k < T; k++) tap[k] = __ldg(taps + k);

    for (int u = warp; u < cnt; u += 8) {
        ring.wait_slot(u);
        const unsigned char *slot_base = smem + (u % C::NS) * C::SLOT_BYTES;
#pragma unroll 1
        for (int p = 0; p < S; p++) {
            const unsigned char *w = slot_base + (p * 32 + lane) * C::LANE_BYTES;
            const long long o0 = (s0 + u) * (long long)C::SLOT_OUT + (p * 32 + lane) * R;   // this lane's first output
            if (CPLX) {
                u64 acc[R];
#pragma unroll
                for (int r = 0; r < R; r++) acc[r] = 0ULL;
#pragma unroll
                for (int c = 0; c < C::NCH; c++) {
                    const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(w)[c];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int k0 = 2 * c - r * D, k1 = 2 * c + 1 - r * D;
                        if (k0 >= 0 && k0 < T) acc[r] = ffma2(v.x, dup2(tap[k0 < 0 ? 0 : (k0 >= T ? 0 : k0)]), acc[r]);
                        if (k1 >= 0 && k1 < T) acc[r] = ffma2(v.y, dup2(tap[k1 < 0 ? 0 : (k1 >= T ? 0 : k1)]), acc[r]);
                    }
                }
                u64 *os = reinterpret_cast<u64 *>(out_v) + o0;
                if (CAN16 && vec_store) {
                    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(os);
#pragma unroll
                    for (int r = 0; r + 1 < R; r += 2) o[r / 2] = make_ulonglong2(acc[r], acc[r + 1]);
                } else {
#pragma unroll
                    for (int r = 0; r < R; r++) os[r] = acc[r];
                }
            } else if (PAIR) {
                u64 acc2[R / 2];
#pragma unroll
                for (int p2 = 0; p2 < R / 2; p2++) acc2[p2] = 0ULL;
                float carry = 0.0f;   // x[4 c4 - 1]
#pragma unroll
                for (int c4 = 0; c4 < C::NCH; c4++) {
                    const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(w)[c4];   // (x0, x1), (x2, x3)
                    float x0, x1, x2, x3;
                    unpack2(v.x, x0, x1);
                    unpack2(v.y, x2, x3);
                    const u64 pr[4] = {pack2(carry, x0), v.x, pack2(x1, x2), v.y};       // pairs starting at 4 c4 - 1, 4 c4, 4 c4 + 1, 4 c4 + 2
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int s0 = 4 * c4 - 1 + i;                                   // index of the pair's first sample
                        if (s0 < 0) continue;
#pragma unroll
                        for (int p2 = 0; p2 < R / 2; p2++) {
                            const int k = s0 - 2 * p2;
                            if (k >= 0 && k < T) acc2[p2] = ffma2(pr[i], dup2(tap[k < 0 ? 0 : (k >= T ? 0 : k)]), acc2[p2]);
                        }
                    }
                    carry = x3;
                }
                float *os = reinterpret_cast<float *>(out_v) + o0;
                if (CAN16 && vec_store) {
                    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(os);
#pragma unroll
                    for (int p2 = 0; p2 + 1 < R / 2; p2 += 2) o[p2 / 2] = make_ulonglong2(acc2[p2], acc2[p2 + 1]);
                } else {
#pragma unroll
                    for (int p2 = 0; p2 < R / 2; p2++) { float a, b; unpack2(acc2[p2], a, b); os[2 * p2] = a; os[2 * p2 + 1] = b; }
                }
            } else {
                float acc[R];
#pragma unroll
                for (int r = 0; r < R; r++) acc[r] = 0.0f;
#pragma unroll
                for (int c4 = 0; c4 < C::NCH; c4++) {
                    const float4 v = reinterpret_cast<const float4 *>(w)[c4];
                    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const int k = 4 * c4 + i - r * D;
                            if (k >= 0 && k < T) acc[r] = fmaf(tap[k < 0 ? 0 : (k >= T ? 0 : k)], e[i], acc[r]);
                        }
                    }
                }
                float *os = reinterpret_cast<float *>(out_v) + o0;
                if (CAN16 && vec_store) {
                    float4 *o = reinterpret_cast<float4 *>(os);
#pragma unroll
                    for (int r = 0; r + 3 < R; r += 4) o[r / 4] = make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
                } else if (CAN8 && vec_store) {
                    float2 *o = reinterpret_cast<float2 *>(os);
#pragma unroll
                    for (int r = 0; r + 1 < R; r += 2) o[r / 2] = make_float2(acc[r], acc[r + 1]);
                } else {   // output not aligned (a pipe's FIFO cursor): scalar stores, L2 merges the sectors
#pragma unroll
                    for (int r = 0; r < R; r++) os[r] = acc[r];
                }
            }
        }
        ring.release_and_refill(u, lane);
    }
}

// Same ring, 2-parallel fast-FIR arithmetic (fir_ffa.cuh): three half-length sub-filters per two outputs instead of
// four -- 17 % fewer FP32-pipe operations per lane pass for a kernel that pipe bounds.  Results agree with the direct
// form to rounding (1.5e-6 of the output scale), not bit for bit, so it is opt-in (sdr_ctx_set_fast_fir, SDR_B200_FIR_FFA=1) until the
// full-size parity properties that rely on ring == generic have a tolerance-based twin.  Real data, stride 1 only.
template <int T, int R, int S>
__global__ void __launch_bounds__(256, 1)
k_fir_r_ffa_ring(const void *__restrict__ in_v, void *__restrict__ out_v, const float *__restrict__ taps, long long n_slots) {
    typedef FirRCfg<false, T, 1, R, S> C;
    const float *in = reinterpret_cast<const float *>(in_v);
    float *out = reinterpret_cast<float *>(out_v);
    const bool vec_store = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long q = n_slots / gridDim.x, rem = n_slots % gridDim.x;
    long long s0 = blockIdx.x * q + (blockIdx.x < rem ? blockIdx.x : rem);
    int cnt = (int)(q + (blockIdx.x < rem ? 1 : 0));
    if (cnt == 0) return;

    typename C::Ring ring;
    ring.init(smem, reinterpret_cast<const unsigned char *>(in + s0 * C::SLOT_OUT), cnt);
    ring.prologue(warp, lane, 8);

    float h0[T / 2], h1[T / 2], hs[T / 2];
#pragma unroll
    for (int j = 0; j < T / 2; j++) {
        h0[j] = __ldg(taps + 2 * j);
        h1[j] = __ldg(taps + 2 * j + 1);
        hs[j] = __fadd_rn(h0[j], h1[j]);
    }

    for (int u = warp; u < cnt; u += 8) {
        ring.wait_slot(u);
        const float *slot_base = reinterpret_cast<const float *>(smem + (u % C::NS) * C::SLOT_BYTES);
        float *out_slot = out + (s0 + u) * C::SLOT_OUT;
#pragma unroll 1
        for (int p = 0; p < S; p++) {
            const float4 *w = reinterpret_cast<const float4 *>(slot_base + (p * 32 + lane) * R);
            float acc[R];
            fir_ffa_lane<T, R>(w, h0, h1, hs, acc);
            float *os = out_slot + (p * 32 + lane) * R;
            if (vec_store) {
                float4 *o = reinterpret_cast<float4 *>(os);
#pragma unroll
                for (int r = 0; r < R; r += 4) o[r / 4] = make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) os[r] = acc[r];
            }
        }
        ring.release_and_refill(u, lane);
    }
}

template <bool CPLX, int T, int D, int R, int S, bool FFA = false, bool PAIR = false>
static int launch_fir_b(Ctx *c, const float *d_taps, const void *d_in, long long n_in, void *d_out, long long num, long long *done) {
    typedef FirRCfg<CPLX, T, D, R, S> C;
    void (*kernel)(const void *, void *, const float *, long long) = k_fir_ring<CPLX, T, D, R, S, PAIR>;
    if constexpr (FFA) kernel = k_fir_r_ffa_ring<T, R, S>;
    long long n_slots = num / C::SLOT_OUT;
    long long by_in = (n_in - C::HALO) / C::SLOT_ELEMS;
    if (by_in < n_slots) n_slots = by_in;
    if (n_slots <= 0) { *done = 0; return SDR_OK; }
    SDR_TRY(c->bind());
    SDR_TRY(ring_attr(c, reinterpret_cast<const void *>(kernel), C::Ring::SMEM_BYTES));
    int sms = c->sm_count - c->reserve_sms;
    if (sms < 1) sms = 1;
    int grid = (int)(n_slots < sms ? n_slots : sms);
    kernel<<<grid, 256, C::Ring::SMEM_BYTES, c->s()>>>(d_in, d_out, d_taps, n_slots);
    c->launches++;
    SDR_CUDA(cudaGetLastError());
    *done = n_slots * C::SLOT_OUT;
    return SDR_OK;
}

// stride 1 and 2, real or complex data; taps_stored = the record's tap count (d_taps zero-padded to >= 128 floats): the
// next larger instantiation serves it.  Same contract as launch_dec_fast; ragged ends are the caller's (generic kernel).
int launch_fir_small_stride_fast(Ctx *c, bool cplx, int taps_stored, int D, const float *d_taps, const void *d_in, long long n_in,
                                 void *d_out, long long num, long long *done, const char **name) {
    *done = 0;
    *name = cplx ? "fir_direct" : "fir_tile";
    if ((D != 1 && D != 2) || (((uintptr_t)d_in) & 15) != 0) return SDR_OK;   // TMA needs a 16-byte aligned source; any output alignment
    const int T = taps_stored <= 32 ? 32 : taps_stored <= 64 ? 64 : taps_stored <= 128 ? 128 : 0;
    if (T == 0) return SDR_OK;
    // a window reaches up to T - taps_stored elements further than the record's own taps need: those reads must stay inside
    // the resident data (they meet zero taps), which launch_fir_b's `by_in` guarantees with the kernel's T
    const bool ffa = c->fir_ffa && !cplx && D == 1 && taps_stored == T;   // sdr_ctx_set_fast_fir / SDR_B200_FIR_FFA
    if (ffa && T == 64) { *name = "fir_r_ffa_ring<64,20,6>"; return launch_fir_b<false, 64, 1, 20, 6, true>(c, d_taps, d_in, n_in, d_out, num, done); }
    if (ffa && T == 32) { *name = "fir_r_ffa_ring<32,20,6>"; return launch_fir_b<false, 32, 1, 20, 6, true>(c, d_taps, d_in, n_in, d_out, num, done); }
    if (ffa && T == 128) { *name = "fir_r_ffa_ring<128,20,6>"; return launch_fir_b<false, 128, 1, 20, 6, true>(c, d_taps, d_in, n_in, d_out, num, done); }
    // 128-tap real stride-1 filter: two adjacent outputs per FFMA2 (same bits, half the FP32 instructions, 181 instead of 197
    // registers): 202.6 vs 178 Gsamples/s.  With 64 / 32 taps the register moves that build the odd-aligned sample pairs cost
    // more than the halved FMA issue saves (379 vs 465, 682 vs 704 Gsamples/s measured): those stay scalar.
    // SDR_B200_FIR_PAIR=0 / =2 force the scalar / the paired form everywhere (measurement knob).
    static const int pair_mode = getenv("SDR_B200_FIR_PAIR") ? atoi(getenv("SDR_B200_FIR_PAIR")) : 1;
    if (pair_mode && !cplx && D == 1) {
        if (T == 128) { *name = "fir_r_ring<128,20,6>"; return launch_fir_b<false, 128, 1, 20, 6, false, true>(c, d_taps, d_in, n_in, d_out, num, done); }
        if (pair_mode == 2 && T == 64) { *name = "fir_r_ring<64,20,6>"; return launch_fir_b<false, 64, 1, 20, 6, false, true>(c, d_taps, d_in, n_in, d_out, num, done); }
        if (pair_mode == 2 && T == 32) { *name = "fir_r_ring<32,20,6>"; return launch_fir_b<false, 32, 1, 20, 6, false, true>(c, d_taps, d_in, n_in, d_out, num, done); }
    }
#define SDR_FIRB(CP, TT, DD, RR, label)                                                     \
    if (cplx == CP && T == TT && D == DD) { *name = label; return launch_fir_b<CP, TT, DD, RR, 6>(c, d_taps, d_in, n_in, d_out, num, done); }
    SDR_FIRB(false, 64, 1, 20, "fir_r_ring<64,20,6>")
    SDR_FIRB(false, 32, 1, 20, "fir_r_ring<32,20,6>")
    SDR_FIRB(false, 128, 1, 20, "fir_r_ring<128,20,6>")
    SDR_FIRB(false, 64, 2, 10, "dec_r_ring<64,2,10>")
    SDR_FIRB(false, 32, 2, 10, "dec_r_ring<32,2,10>")
    SDR_FIRB(false, 128, 2, 10, "dec_r_ring<128,2,10>")
    SDR_FIRB(true, 64, 1, 10, "fir_c_ring<64,10,6>")
    SDR_FIRB(true, 32, 1, 10, "fir_c_ring<32,10,6>")
    SDR_FIRB(true, 128, 1, 10, "fir_c_ring<128,10,6>")
    SDR_FIRB(true, 64, 2, 5, "dec_c_ring<64,2,5>")
    SDR_FIRB(true, 32, 2, 5, "dec_c_ring<32,2,5>")
    SDR_FIRB(true, 128, 2, 5, "dec_c_ring<128,2,5>")
#undef SDR_FIRB
    return SDR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// real rational resampler L/M (gcd 1): output k uses taps c[f_k + l L], window start ceil(k M / L)
// ---------------------------------------------------------------------------------------------------------------
template <int L, int M, int T>
struct Phase {
    static constexpr __host__ __device__ int f(int j) { return (L - (j * M) % L) % L; }          // first tap of phase j
    static constexpr __host__ __device__ int i0(int j) { return (j * M + L - 1) / L; }           // window start in the cycle
    static constexpr __host__ __device__ int len(int j) { return (T - f(j) + L - 1) / L; }       // taps of phase j
    static constexpr __host__ __device__ int max_end() {
        int m = 0;
        for (int j = 0; j < L; j++) { int e = i0(j) + len(j); if (e > m) m = e; }
        return m;
    }
};

template <bool CPLX, int L, int M, int T, int CY, int S>
struct ResRCfg {
    static constexpr int EB = CPLX ? 8 : 4;
    static constexpr int EPC = 16 / EB;
    static constexpr int LANE_IN = CY * M;                   // input elements a lane advances per pass
    static constexpr int LANE_OUT = CY * L;
    static constexpr int LANE_BYTES = LANE_IN * EB;
    static_assert(LANE_BYTES % 16 == 0 && (LANE_BYTES / 16) % 2 == 1, "lane stride must be an odd number of 16-byte chunks");
    static_assert(CPLX || LANE_OUT % 2 == 0, "real outputs are stored in 8-byte pairs");
    static constexpr int WIN = (CY - 1) * M + Phase<L, M, T>::max_end();
    static constexpr int NCH = (WIN + EPC - 1) / EPC;
    static constexpr int HALO = (NCH * EPC > LANE_IN) ? NCH * EPC - LANE_IN : EPC;
    static constexpr int HALO_BYTES = HALO * EB;
    static constexpr int SLOT_IN = 32 * LANE_IN * S;
    static constexpr int SLOT_OUT = 32 * LANE_OUT * S;
    static constexpr int SLOT_BYTES = SLOT_IN * EB;
    static constexpr int NS = (220 * 1024 - HALO_BYTES - 512) / SLOT_BYTES;
    typedef ContigRing<SLOT_BYTES, HALO_BYTES, NS> Ring;
    static_assert(NS >= 11, "ring too small for 8 warps plus prefetch");
};

// CPLX: complex data, real taps (resampleAVXRC, resample.c:125-142): the same walk with one FFMA2 per (sample, tap)
template <bool CPLX, int L, int M, int T, int CY, int S>
__global__ void __launch_bounds__(256, 1)
k_res_ring(const void *__restrict__ in_v, void *__restrict__ out_v, const float *__restrict__ taps, long long n_slots) {
    typedef ResRCfg<CPLX, L, M, T, CY, S> C;
    const bool vec_store = (reinterpret_cast<uintptr_t>(out_v) & 7) == 0;
    typedef Phase<L, M, T> P;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long q = n_slots / gridDim.x, rem = n_slots % gridDim.x;
    long long s0 = blockIdx.x * q + (blockIdx.x < rem ? blockIdx.x : rem);
    int cnt = (int)(q + (blockIdx.x < rem ? 1 : 0));
    if (cnt == 0) return;

    typename C::Ring ring;
    ring.init(smem, reinterpret_cast<const unsigned char *>(in_v) + s0 * (long long)C::SLOT_BYTES, cnt);
    ring.prologue(warp, lane, 8);

    float tap[T];   // the plain tap list; phase j, tap l is tap[f(j) + l * L]
#pragma unroll
    for (int k = 0; k < T; k++) tap[k] = __ldg(taps + k);

    for (int u = warp; u < cnt; u += 8) {
        ring.wait_slot(u);
        const unsigned char *slot_base = smem + (u % C::NS) * C::SLOT_BYTES;
#pragma unroll 1
        for (int p = 0; p < S; p++) {
            const unsigned char *w = slot_base + (p * 32 + lane) * C::LANE_BYTES;
            const long long o0 = (s0 + u) * (long long)C::SLOT_OUT + (p * 32 + lane) * C::LANE_OUT;
            if (CPLX) {
                u64 acc[C::LANE_OUT];
#pragma unroll
                for (int o = 0; o < C::LANE_OUT; o++) acc[o] = 0ULL;
#pragma unroll
                for (int c = 0; c < C::NCH; c++) {
                    const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(w)[c];
                    const u64 e[2] = {v.x, v.y};
#pragma unroll
                    for (int i = 0; i < 2; i++) {
#pragma unroll
                        for (int cy = 0; cy < CY; cy++) {
#pragma unroll
                            for (int j = 0; j < L; j++) {
                                const int l = 2 * c + i - cy * M - P::i0(j);
                                if (l >= 0 && l < P::len(j)) acc[cy * L + j] = ffma2(e[i], dup2(tap[(l >= 0 && l < P::len(j)) ? P::f(j) + l * L : 0]), acc[cy * L + j]);
                            }
                        }
                    }
                }
                u64 *os = reinterpret_cast<u64 *>(out_v) + o0;
#pragma unroll
                for (int o = 0; o < C::LANE_OUT; o++) os[o] = acc[o];
            } else {
                float acc[C::LANE_OUT];
#pragma unroll
                for (int o = 0; o < C::LANE_OUT; o++) acc[o] = 0.0f;
#pragma unroll
                for (int c4 = 0; c4 < C::NCH; c4++) {
                    const float4 v = reinterpret_cast<const float4 *>(w)[c4];
                    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
#pragma unroll
                        for (int cy = 0; cy < CY; cy++) {
#pragma unroll
                            for (int j = 0; j < L; j++) {
                                const int l = 4 * c4 + i - cy * M - P::i0(j);
                                if (l >= 0 && l < P::len(j)) acc[cy * L + j] = fmaf(tap[(l >= 0 && l < P::len(j)) ? P::f(j) + l * L : 0], e[i], acc[cy * L + j]);
                            }
                        }
                    }
                }
                float *os = reinterpret_cast<float *>(out_v) + o0;
                if (vec_store) {
                    float2 *o2 = reinterpret_cast<float2 *>(os);
#pragma unroll
                    for (int o = 0; o < C::LANE_OUT; o += 2) o2[o / 2] = make_float2(acc[o], acc[o + 1]);
                } else {
#pragma unroll
                    for (int o = 0; o < C::LANE_OUT; o++) os[o] = acc[o];
                }
            }
        }
        ring.release_and_refill(u, lane);
    }
}

template <bool CPLX, int L, int M, int T, int CY, int S>
static int launch_res(Ctx *c, const float *d_taps, const void *d_in, long long n_in, void *d_out, long long num, long long *done) {
    typedef ResRCfg<CPLX, L, M, T, CY, S> C;
    long long n_slots = num / C::SLOT_OUT;
    long long by_in = (n_in - C::HALO) / C::SLOT_IN;
    if (by_in < n_slots) n_slots = by_in;
    if (n_slots <= 0) { *done = 0; return SDR_OK; }
    SDR_TRY(c->bind());
    SDR_TRY(ring_attr(c, reinterpret_cast<const void *>(k_res_ring<CPLX, L, M, T, CY, S>), C::Ring::SMEM_BYTES));
    int sms = c->sm_count - c->reserve_sms;
    if (sms < 1) sms = 1;
    int grid = (int)(n_slots < sms ? n_slots : sms);
    k_res_ring<CPLX, L, M, T, CY, S><<<grid, 256, C::Ring::SMEM_BYTES, c->s()>>>(d_in, d_out, d_taps, n_slots);
    c->launches++;
    SDR_CUDA(cudaGetLastError());
    *done = n_slots * C::SLOT_OUT;
    return SDR_OK;
}

// d_plain_taps: the n_taps plain coefficients on the device.  The window of output 0 must start at d_in (phase 0).
int launch_res_fast(Ctx *c, bool cplx, int L, int M, int n_taps, const float *d_plain_taps, const void *d_in, long long n_in,
                    void *d_out, long long num, long long *done, const char **name) {
    *done = 0;
    *name = "fir_tile";
    if ((((uintptr_t)d_in) & 15) != 0) return SDR_OK;
    // passes per ring slot (S): smaller slots = more, finer-grained copies in flight.  Measurement knob; every S gives
    // the same bits (a lane's arithmetic does not depend on how lanes are grouped into slots).
    static const int res_s = getenv("SDR_B200_RES_S") ? atoi(getenv("SDR_B200_RES_S")) : 2;
    if (!cplx && L == 3 && M == 10 && n_taps == 90 && res_s == 1) {
        *name = "res_r_ring<3,10,90,6,1>";
        return launch_res<false, 3, 10, 90, 6, 1>(c, d_plain_taps, d_in, n_in, d_out, num, done);
    }
    if (!cplx && L == 3 && M == 10 && n_taps == 90) {
        *name = "res_r_ring<3,10,90,6,2>";
        return launch_res<false, 3, 10, 90, 6, 2>(c, d_plain_taps, d_in, n_in, d_out, num, done);
    }
    if (!cplx && L == 3 && M == 10 && n_taps == 31) {   // examples/fm/Coeffs.hs:76-110, the FM receiver's own audio resampler
        *name = "res_r_ring<3,10,31,6,2>";
        return launch_res<false, 3, 10, 31, 6, 2>(c, d_plain_taps, d_in, n_in, d_out, num, done);
    }
    // complex data (fastResamplerC -> resampleAVXRC): a lane owns 3 cycles (9 outputs from 30 inputs, lane stride 240 B)
    if (cplx && L == 3 && M == 10 && n_taps == 90) {
        *name = "res_c_ring<3,10,90,3,2>";
        return launch_res<true, 3, 10, 90, 3, 2>(c, d_plain_taps, d_in, n_in, d_out, num, done);
    }
    if (cplx && L == 3 && M == 10 && n_taps == 31) {
        *name = "res_c_ring<3,10,31,3,2>";
        return launch_res<true, 3, 10, 31, 3, 2>(c, d_plain_taps, d_in, n_in, d_out, num, done);
    }
    return SDR_OK;
}

}  // namespace sdr
