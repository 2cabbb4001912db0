// kernels_lowrate.cu -- fused low-rate end of the FM receiver chain (cfg4, reference examples/fm/fm.hs:38-40):
//
//     firResampler resp samples >-> firFilter filt samples >-> P.map (VG.map (* 0.2))
//
// as ONE kernel: float phases in, audio samples out, the resampled stream never touches HBM.  Un-fused, each push of the
// chain costs three tuned launches, their generic prefixes / ragged ends and an element-wise launch behind the front end,
// plus the copies between the stages' buffers -- fixed cost that dominated the chain in round 1 (profiles/r01_launches_fm_chain.csv).
//
// Flat-stream semantics:   r[k] = sum_l cr[f_k + l L] x[i_k + l],  f_k = (-k M) mod L,  i_k = ceil(k M / L)
//                                                 z[n] = scale * sum_t cf[t] r[n + t]
// A CTA produces a tile of NZ = 1470 outputs z.  Phase 1: the 256 threads stage the tile's input span in shared memory
// (coalesced) and each computes 2 whole resampler cycles (6 outputs r from 20 inputs, compile-time phase pattern, taps in
// registers, increasing tap order) into shared memory: 1536 values r, of which the tile needs NZ + TF - 1 = 1533.  Phase 2:
// each thread computes 4 consecutive outputs z from a 67-value window of r (17 LDS.128 per 256 FFMA, filter taps in
// registers, increasing tap order), multiplies by `scale` with one rounding and stores.  Every rounding of the un-fused
// chain is reproduced (r is rounded to binary32 when it is written to shared memory, as it is when the resampler stage
// writes it to HBM), so the fused stage is bit-identical to the three stages one after the other.
// Both tap sets travel as LAUNCH PARAMETERS (616 bytes, by value): with every index a compile-time constant each tap is a
// constant-bank operand of its FFMA -- no tap registers, no tap loads per tile (the first version re-read 154 taps per
// thread and tile into registers, 128 registers per thread, 2 CTAs per SM; the kernel is a chain of short dependent
// phases -- stage, resample, filter -- and what hides their latency is other CTAs on the same SM).
// Roofline: 4 B in + 1.2 B out per input sample and 9 + 19.2 FMA per input sample: FP32-pipe bound, but the whole
// low-rate end is ~11 % of the chain's arithmetic.
#include "ring_common.cuh"

#include <cstdlib>

namespace sdr {

#ifndef LOWRATE_MIN_CTAS
#define LOWRATE_MIN_CTAS 5
#endif

template <int L, int M, int T>
struct LrPhase {
    static constexpr __host__ __device__ int f(int j) { return (L - (j * M) % L) % L; }
    static constexpr __host__ __device__ int i0(int j) { return (j * M + L - 1) / L; }
    static constexpr __host__ __device__ int len(int j) { return (T - f(j) + L - 1) / L; }
    static constexpr __host__ __device__ int max_end() {
        int m = 0;
        for (int j = 0; j < L; j++) { int e = i0(j) + len(j); if (e > m) m = e; }
        return m;
    }
};

template <int L, int M, int TR, int TF>
struct LowCfg {
    static constexpr int CY = 2;                              // resampler cycles per thread
    static constexpr int NT = 256;
    static constexpr int LANE_IN = CY * M;                    // 20 floats = 80 B: odd number of 16-byte chunks
    static_assert((LANE_IN * 4) % 16 == 0 && ((LANE_IN * 4) / 16) % 2 == 1, "thread stride must be an odd number of 16-byte chunks");
    static constexpr int NR = NT * CY * L;                    // 1536 resampler outputs per tile
    static constexpr int NZ = ((NR - (TF - 1) - (L - 1)) / (4 * L)) * (4 * L) - ((((NR - (TF - 1) - (L - 1)) / (4 * L)) * (4 * L)) % L);
    static constexpr int WIN = (CY - 1) * M + LrPhase<L, M, TR>::max_end();
    static constexpr int NCH = (WIN + 3) / 4;
    static constexpr int XS = (NT - 1) * LANE_IN + NCH * 4;   // staged input floats
    static constexpr int SMEM_BYTES = (XS + NR + 8) * 4;
    static_assert(NZ % L == 0 && NZ + TF - 1 + L - 1 <= NR, "tile geometry");
};

__device__ __forceinline__ float seg_load(const float *a, long long na, const float *b, long long nb, long long i) {
    if (i < 0) return 0.0f;
    if (i < na) return __ldg(a + i);
    i -= na;
    return i < nb ? __ldg(b + i) : 0.0f;
}

template <int TR, int TF> struct LowTaps { float r[TR]; float f[TF]; };

template <int L, int M, int TR, int TF>
__global__ void __launch_bounds__(256, LOWRATE_MIN_CTAS)
k_fm_lowrate(const float *__restrict__ xa, long long na, const float *__restrict__ xb, long long nb, long long x0_global,
             long long n0, float *__restrict__ out, long long num, const __grid_constant__ LowTaps<TR, TF> K, float scale) {
    typedef LowCfg<L, M, TR, TF> C;
    typedef LrPhase<L, M, TR> P;
    extern __shared__ __align__(16) float sm[];
    float *xs = sm;                    // [XS]
    float *rs = sm + ((C::XS + 3) / 4) * 4;   // [NR]
    const int t = threadIdx.x;
    // programmatic dependent launch (see k_dec_ring): the next launch may be scheduled while this grid drains; this one
    // touches global memory only after everything before it in the stream has completed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    for (long long tile = blockIdx.x; tile * C::NZ < num; tile += gridDim.x) {
        const long long n_tile = n0 + tile * C::NZ;            // first output z of the tile (global index)
        const int off = (int)(n_tile % L);                     // resampler outputs before it in its cycle
        const long long c0 = n_tile / L;                       // first cycle computed
        const long long x_first = c0 * M - x0_global;          // index into x (= xa ++ xb) of that cycle's first sample
        __syncthreads();                                       // previous tile's phase 2 has finished with rs / xs
        {   // all of a thread's loads are issued before the first one is used (21 independent loads in flight)
            constexpr int PER = (C::XS + C::NT - 1) / C::NT;
            float v[PER];
            const bool inside_a = x_first >= 0 && x_first + C::XS <= na;   // the common case: no boundary, no bounds
#pragma unroll
            for (int j = 0; j < PER; j++) {
                const int i = t + j * C::NT;
                v[j] = i < C::XS ? (inside_a ? __ldg(xa + x_first + i) : seg_load(xa, na, xb, nb, x_first + i)) : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < PER; j++) {
                const int i = t + j * C::NT;
                if (i < C::XS) xs[i] = v[j];
            }
        }
        __syncthreads();
        {   // phase 1: thread t -> cycles c0 + 2t, c0 + 2t + 1
            const float4 *w = reinterpret_cast<const float4 *>(xs + t * C::LANE_IN);
            float acc[C::CY * L];
#pragma unroll
            for (int o = 0; o < C::CY * L; o++) acc[o] = 0.0f;
#pragma unroll
            for (int c4 = 0; c4 < C::NCH; c4++) {
                const float4 v = w[c4];
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
#pragma unroll
                    for (int cy = 0; cy < C::CY; cy++) {
#pragma unroll
                        for (int j = 0; j < L; j++) {
                            const int l = 4 * c4 + i - cy * M - P::i0(j);
                            if (l >= 0 && l < P::len(j))
                                acc[cy * L + j] = fmaf(K.r[(l >= 0 && l < P::len(j)) ? P::f(j) + l * L : 0], e[i], acc[cy * L + j]);
                        }
                    }
                }
            }
            // r[c0 L + 6 t + o] lands at rs[6 t + o - off]: index 0 is r[n_tile]
#pragma unroll
            for (int o = 0; o < C::CY * L; o++) {
                const int idx = t * C::CY * L + o - off;
                if (idx >= 0) rs[idx] = acc[o];
            }
        }
        __syncthreads();
        {   // phase 2: 4 consecutive outputs per thread and pass
            const long long left = num - tile * C::NZ;
            const int nz = (int)(left < C::NZ ? left : C::NZ);
            for (int g = t; 4 * g < nz; g += C::NT) {
                const float4 *w = reinterpret_cast<const float4 *>(rs + 4 * g);
                float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int c4 = 0; c4 < (TF + 3 + 3) / 4; c4++) {
                    const float4 v = w[c4];
                    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const int k = 4 * c4 + i - r;
                            if (k >= 0 && k < TF) acc[r] = fmaf(K.f[(k >= 0 && k < TF) ? k : 0], e[i], acc[r]);
                        }
                    }
                }
                float *os = out + tile * C::NZ + 4 * g;
                const float z[4] = {__fmul_rn(acc[0], scale), __fmul_rn(acc[1], scale), __fmul_rn(acc[2], scale), __fmul_rn(acc[3], scale)};
                if (4 * g + 4 <= nz && (reinterpret_cast<uintptr_t>(os) & 15) == 0) {
                    *reinterpret_cast<float4 *>(os) = make_float4(z[0], z[1], z[2], z[3]);
                } else {
#pragma unroll
                    for (int r = 0; r < 4; r++) if (4 * g + r < nz) os[r] = z[r];
                }
            }
        }
    }
}

template <int L, int M, int TR, int TF>
static int launch_low(Ctx *c, const float *h_taps_r, const float *h_taps_f, float scale, Seg2 seg, long long n0, float *d_out,
                      long long num) {
    typedef LowCfg<L, M, TR, TF> C;
    SDR_TRY(c->bind());
    SDR_TRY(ring_attr(c, reinterpret_cast<const void *>(k_fm_lowrate<L, M, TR, TF>), C::SMEM_BYTES));
    long long tiles = (num + C::NZ - 1) / C::NZ;
    long long cap = 2LL * c->sm_count * 16;
    int grid = (int)(tiles < cap ? tiles : cap);
    const long long x0_global = (n0 * M + L - 1) / L;
    LowTaps<TR, TF> K;
    for (int k = 0; k < TR; k++) K.r[k] = h_taps_r[k];
    for (int k = 0; k < TF; k++) K.f[k] = h_taps_f[k];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = c->s();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("SDR_B200_NO_PDL") != nullptr;   // measurement knob
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    SDR_CUDA(cudaLaunchKernelEx(&cfg, k_fm_lowrate<L, M, TR, TF>, (const float *)seg.a, seg.na, (const float *)seg.b, seg.nb, x0_global,
                                n0, d_out, num, K, scale));
    c->launches++;
    SDR_CUDA(cudaGetLastError());
    return SDR_OK;
}

int launch_fm_lowrate(Ctx *c, int L, int M, int n_taps_r, const float *h_taps_r, int n_taps_f, const float *h_taps_f, float scale,
                      Seg2 seg, long long n0, float *d_out, long long num, long long *done, const char **name) {
    *done = 0;
    *name = "unfused";
    if (num <= 0) return SDR_OK;
    if (L == 3 && M == 10 && n_taps_f == 64 && n_taps_r == 90) {
        *name = "fm_lowrate<3,10,90,64>";
        SDR_TRY((launch_low<3, 10, 90, 64>(c, h_taps_r, h_taps_f, scale, seg, n0, d_out, num)));
        *done = num;
    } else if (L == 3 && M == 10 && n_taps_f == 64 && n_taps_r == 31) {   // the FM example's own sets (examples/fm/Coeffs.hs)
        *name = "fm_lowrate<3,10,31,64>";
        SDR_TRY((launch_low<3, 10, 31, 64>(c, h_taps_r, h_taps_f, scale, seg, n0, d_out, num)));
        *done = num;
    }
    return SDR_OK;
}

}  // namespace sdr
