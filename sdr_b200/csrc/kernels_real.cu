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
// real FIR, stride 1
// ---------------------------------------------------------------------------------------------------------------
template <int T, int R, int S>
struct FirRCfg {
    static_assert(R % 4 == 0 && (R / 4) % 2 == 1, "lane stride must be an odd number of 16-byte chunks");
    static constexpr int PASS_OUT = 32 * R;
    static constexpr int SLOT_OUT = PASS_OUT * S;            // outputs == input floats per slot
    static constexpr int SLOT_BYTES = SLOT_OUT * 4;
    static constexpr int WIN4 = (R + T - 1 + 3) / 4;         // LDS.128 per lane pass
    static constexpr int HALO = WIN4 * 4 - R;                // floats a pass reads beyond its own outputs' positions
    static constexpr int HALO_BYTES = HALO * 4;
    static constexpr int NS = (220 * 1024 - HALO_BYTES - 512) / SLOT_BYTES;
    typedef ContigRing<SLOT_BYTES, HALO_BYTES, NS> Ring;
    static_assert(NS >= 11, "ring too small for 8 warps plus prefetch");
};

template <int T, int R, int S>
__global__ void __launch_bounds__(256, 1)
k_fir_r_ring(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ taps, long long n_slots) {
    typedef FirRCfg<T, R, S> C;
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

    float tap[T];
#pragma unroll
    for (int k = 0; k < T; k++) tap[k] = __ldg(taps + k);

    for (int u = warp; u < cnt; u += 8) {
        ring.wait_slot(u);
        const float *slot_base = reinterpret_cast<const float *>(smem + (u % C::NS) * C::SLOT_BYTES);
        float *out_slot = out + (s0 + u) * C::SLOT_OUT;
#pragma unroll 1
        for (int p = 0; p < S; p++) {
            const float4 *w = reinterpret_cast<const float4 *>(slot_base + (p * 32 + lane) * R);
            float acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = 0.0f;
#pragma unroll
            for (int c4 = 0; c4 < C::WIN4; c4++) {
                const float4 v = w[c4];
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int k = 4 * c4 + i - r;
                        if (k >= 0 && k < T) acc[r] = fmaf(tap[k], e[i], acc[r]);
                    }
                }
            }
            float *os = out_slot + (p * 32 + lane) * R;
            if (vec_store) {
                float4 *o = reinterpret_cast<float4 *>(os);
#pragma unroll
                for (int r = 0; r < R; r += 4) o[r / 4] = make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
            } else {   // output not 16-byte aligned (a pipe's FIFO cursor): scalar stores, L2 merges the sectors
#pragma unroll
                for (int r = 0; r < R; r++) os[r] = acc[r];
            }
        }
        ring.release_and_refill(u, lane);
    }
}

// Same ring, 2-parallel fast-FIR arithmetic (fir_ffa.cuh): three half-length sub-filters per two outputs instead of
// four -- 17 % fewer FP32-pipe operations per lane pass for a kernel that pipe bounds.  Results agree with the direct
// form to rounding (1.5e-6 of the output scale), not bit for bit, so it is opt-in (sdr_ctx_set_fast_fir, SDR_B200_FIR_FFA=1) until the
// full-size parity properties that rely on ring == generic have a tolerance-based twin.
template <int T, int R, int S>
__global__ void __launch_bounds__(256, 1)
k_fir_r_ffa_ring(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ taps, long long n_slots) {
    typedef FirRCfg<T, R, S> C;
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

template <int T, int R, int S, bool FFA = false>
static int launch_fir_r(Ctx *c, const float *d_taps, const float *d_in, long long n_in, float *d_out, long long num,
                        long long *done) {
    typedef FirRCfg<T, R, S> C;
    auto kernel = FFA ? k_fir_r_ffa_ring<T, R, S> : k_fir_r_ring<T, R, S>;
    long long n_slots = num / C::SLOT_OUT;
    long long by_in = (n_in - C::HALO) / C::SLOT_OUT;
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

int launch_fir_r_fast(Ctx *c, int T, int D, const float *d_taps, const float *d_in, long long n_in, float *d_out,
                      long long num, long long *done, const char **name) {
    *done = 0;
    *name = "fir_tile";
    if (D != 1 || (((uintptr_t)d_in) & 15) != 0) return SDR_OK;   // TMA needs a 16-byte aligned source; any output alignment
    const bool ffa = c->fir_ffa;   // sdr_ctx_set_fast_fir / SDR_B200_FIR_FFA
    if (T == 64 && ffa) { *name = "fir_r_ffa_ring<64,20,6>"; return launch_fir_r<64, 20, 6, true>(c, d_taps, d_in, n_in, d_out, num, done); }
    if (T == 32 && ffa) { *name = "fir_r_ffa_ring<32,20,6>"; return launch_fir_r<32, 20, 6, true>(c, d_taps, d_in, n_in, d_out, num, done); }
    if (T == 128 && ffa) { *name = "fir_r_ffa_ring<128,20,6>"; return launch_fir_r<128, 20, 6, true>(c, d_taps, d_in, n_in, d_out, num, done); }
    if (T == 64) { *name = "fir_r_ring<64,20,6>"; return launch_fir_r<64, 20, 6>(c, d_taps, d_in, n_in, d_out, num, done); }
    if (T == 32) { *name = "fir_r_ring<32,20,6>"; return launch_fir_r<32, 20, 6>(c, d_taps, d_in, n_in, d_out, num, done); }
    if (T == 128) { *name = "fir_r_ring<128,20,6>"; return launch_fir_r<128, 20, 6>(c, d_taps, d_in, n_in, d_out, num, done); }
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

template <int L, int M, int T, int CY, int S>
struct ResRCfg {
    static constexpr int LANE_IN = CY * M;                   // input floats a lane advances per pass
    static constexpr int LANE_OUT = CY * L;
    static_assert(LANE_IN % 4 == 0 && (LANE_IN / 4) % 2 == 1, "lane stride must be an odd number of 16-byte chunks");
    static_assert(LANE_OUT % 2 == 0, "outputs are stored in 8-byte pairs");
    static constexpr int WIN = (CY - 1) * M + Phase<L, M, T>::max_end();
    static constexpr int WIN4 = (WIN + 3) / 4;
    static constexpr int HALO = (WIN4 * 4 > LANE_IN) ? WIN4 * 4 - LANE_IN : 4;
    static constexpr int HALO_BYTES = HALO * 4;
    static constexpr int SLOT_IN = 32 * LANE_IN * S;
    static constexpr int SLOT_OUT = 32 * LANE_OUT * S;
    static constexpr int SLOT_BYTES = SLOT_IN * 4;
    static constexpr int NS = (220 * 1024 - HALO_BYTES - 512) / SLOT_BYTES;
    typedef ContigRing<SLOT_BYTES, HALO_BYTES, NS> Ring;
    static_assert(NS >= 11, "ring too small for 8 warps plus prefetch");
};

template <int L, int M, int T, int CY, int S>
__global__ void __launch_bounds__(256, 1)
k_res_r_ring(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ taps, long long n_slots) {
    typedef ResRCfg<L, M, T, CY, S> C;
    const bool vec_store = (reinterpret_cast<uintptr_t>(out) & 7) == 0;
    typedef Phase<L, M, T> P;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long q = n_slots / gridDim.x, rem = n_slots % gridDim.x;
    long long s0 = blockIdx.x * q + (blockIdx.x < rem ? blockIdx.x : rem);
    int cnt = (int)(q + (blockIdx.x < rem ? 1 : 0));
    if (cnt == 0) return;

    typename C::Ring ring;
    ring.init(smem, reinterpret_cast<const unsigned char *>(in + s0 * C::SLOT_IN), cnt);
    ring.prologue(warp, lane, 8);

    float tap[T];   // the plain tap list; phase j, tap l is tap[f(j) + l * L]
#pragma unroll
    for (int k = 0; k < T; k++) tap[k] = __ldg(taps + k);

    for (int u = warp; u < cnt; u += 8) {
        ring.wait_slot(u);
        const float *slot_base = reinterpret_cast<const float *>(smem + (u % C::NS) * C::SLOT_BYTES);
        float *out_slot = out + (s0 + u) * C::SLOT_OUT;
#pragma unroll 1
        for (int p = 0; p < S; p++) {
            const float4 *w = reinterpret_cast<const float4 *>(slot_base + (p * 32 + lane) * C::LANE_IN);
            float acc[C::LANE_OUT];
#pragma unroll
            for (int o = 0; o < C::LANE_OUT; o++) acc[o] = 0.0f;
#pragma unroll
            for (int c4 = 0; c4 < C::WIN4; c4++) {
                const float4 v = w[c4];
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
#pragma unroll
                    for (int cy = 0; cy < CY; cy++) {
#pragma unroll
                        for (int j = 0; j < L; j++) {
                            const int l = 4 * c4 + i - cy * M - P::i0(j);
                            if (l >= 0 && l < P::len(j)) acc[cy * L + j] = fmaf(tap[P::f(j) + l * L], e[i], acc[cy * L + j]);
                        }
                    }
                }
            }
            float *os = out_slot + (p * 32 + lane) * C::LANE_OUT;
            if (vec_store) {
                float2 *o2 = reinterpret_cast<float2 *>(os);
#pragma unroll
                for (int o = 0; o < C::LANE_OUT; o += 2) o2[o / 2] = make_float2(acc[o], acc[o + 1]);
            } else {
#pragma unroll
                for (int o = 0; o < C::LANE_OUT; o++) os[o] = acc[o];
            }
        }
        ring.release_and_refill(u, lane);
    }
}

template <int L, int M, int T, int CY, int S>
static int launch_res_r(Ctx *c, const float *d_taps, const float *d_in, long long n_in, float *d_out, long long num,
                        long long *done) {
    typedef ResRCfg<L, M, T, CY, S> C;
    long long n_slots = num / C::SLOT_OUT;
    long long by_in = (n_in - C::HALO) / C::SLOT_IN;
    if (by_in < n_slots) n_slots = by_in;
    if (n_slots <= 0) { *done = 0; return SDR_OK; }
    SDR_TRY(c->bind());
    SDR_TRY(ring_attr(c, reinterpret_cast<const void *>(k_res_r_ring<L, M, T, CY, S>), C::Ring::SMEM_BYTES));
    int sms = c->sm_count - c->reserve_sms;
    if (sms < 1) sms = 1;
    int grid = (int)(n_slots < sms ? n_slots : sms);
    k_res_r_ring<L, M, T, CY, S><<<grid, 256, C::Ring::SMEM_BYTES, c->s()>>>(d_in, d_out, d_taps, n_slots);
    c->launches++;
    SDR_CUDA(cudaGetLastError());
    *done = n_slots * C::SLOT_OUT;
    return SDR_OK;
}

// d_plain_taps: the n_taps plain coefficients on the device.  The window of output 0 must start at d_in (phase 0).
int launch_res_r_fast(Ctx *c, int L, int M, int n_taps, const float *d_plain_taps, const float *d_in, long long n_in,
                      float *d_out, long long num, long long *done, const char **name) {
    *done = 0;
    *name = "fir_tile";
    if ((((uintptr_t)d_in) & 15) != 0) return SDR_OK;
    // passes per ring slot (S): smaller slots = more, finer-grained copies in flight.  Measurement knob; every S gives
    // the same bits (a lane's arithmetic does not depend on how lanes are grouped into slots).
    static const int res_s = getenv("SDR_B200_RES_S") ? atoi(getenv("SDR_B200_RES_S")) : 2;
    if (L == 3 && M == 10 && n_taps == 90 && res_s == 1) {
        *name = "res_r_ring<3,10,90,6,1>";
        return launch_res_r<3, 10, 90, 6, 1>(c, d_plain_taps, d_in, n_in, d_out, num, done);
    }
    if (L == 3 && M == 10 && n_taps == 90) {
        *name = "res_r_ring<3,10,90,6,2>";
        return launch_res_r<3, 10, 90, 6, 2>(c, d_plain_taps, d_in, n_in, d_out, num, done);
    }
    if (L == 3 && M == 10 && n_taps == 31) {   // examples/fm/Coeffs.hs:76-110, the FM receiver's own audio resampler
        *name = "res_r_ring<3,10,31,6,2>";
        return launch_res_r<3, 10, 31, 6, 2>(c, d_plain_taps, d_in, n_in, d_out, num, done);
    }
    return SDR_OK;
}

}  // namespace sdr
