// kernels_generic.cu -- shape-agnostic sm_100a kernels of libsdr_b200: every tap count / factor / alignment, real or
// complex data, input optionally split over two device segments (lastBuf ++ nextBuf).  The tuned kernels in
// kernels_fast.cu take the headline shapes; these finish ragged tails and cover everything else.
//
// Math (reference c_sources/, flat-stream form -- SURVEY.md section 8a):
//   FIR / decimator   y[m] = sum_k c[k] x[m*D + k]                       filter.c:16-22, decimate.c:16-22
//   resampler         y[i] = sum_l g[grp(i)][l] x[start(i) + l]          resample.c:34-49 (group tables)
//   convert u8        (float(b) - 128) * (1/128)                         convert.c:15-20
//   convert i16       float(v) * (1/2048)                                convert.c:52-57
//   convert tx        clamp(int16((v + 1) * 2048) - 2048)                convert.c:87-101
//   scale             in * k                                             scale.c:15-20
//   fm demod          phase(s[n] * conj(s[n-1]))                         Demod.hs:21-36
//   dc blocker        y[n] = x[n] - x[n-1] + 0.997 y[n-1]                filter.c:152-161
#include "common.cuh"
#include "demod.cuh"

#include <cstdlib>
#include <type_traits>

namespace sdr {

static inline int grid_for(long long n, int block, int sm_count, int per_sm = 16) {
    long long g = (n + block - 1) / block;
    long long cap = (long long)sm_count * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

#define SDR_LAUNCH_CHECK(c)                       \
    do {                                          \
        (c)->launches++;                          \
        SDR_CUDA(cudaGetLastError());             \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// input accessor over (a ++ b); elements past the end read as zero
// ---------------------------------------------------------------------------------------------------------------
template <typename E> struct Zero;
template <> struct Zero<float>  { static __device__ __forceinline__ float  v() { return 0.0f; } };
template <> struct Zero<float2> { static __device__ __forceinline__ float2 v() { return make_float2(0.0f, 0.0f); } };

template <typename E>
__device__ __forceinline__ E seg_at(const E *__restrict__ a, long long na, const E *__restrict__ b, long long nb,
                                    long long e) {
    if (e < na) return __ldg(a + e);
    e -= na;
    if (e < nb) return __ldg(b + e);
    return Zero<E>::v();
}

// where output o starts in the input and which tap row it uses
struct OutMap {
    int D;                 // decimation (FIR mode, ng == 0)
    int g0, ng, sum_inc;   // resampler mode when ng > 0
    const int *prefix;     // [ng] exclusive prefix sums of the per-group increments
    int row_stride;        // floats between tap rows
};

__device__ __forceinline__ void map_output(const OutMap &m, long long o, long long *start, int *row) {
    if (m.ng == 0) { *start = o * m.D; *row = 0; return; }
    long long gi = (long long)m.g0 + o;
    long long cyc = gi / m.ng;
    int g = (int)(gi - cyc * m.ng);
    *start = cyc * m.sum_inc + __ldg(m.prefix + g) - __ldg(m.prefix + m.g0);
    *row = g;
}

// ---------------------------------------------------------------------------------------------------------------
// FAST arithmetic: one thread per output, fused multiply-add, taps in increasing order (the same order the tuned
// kernels use, so tuned + generic outputs of one call are computed identically).
// ---------------------------------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256) k_fir_direct(OutMap m, int T, const float *__restrict__ taps,
                                                    const void *__restrict__ a, long long na,
                                                    const void *__restrict__ b, long long nb,
                                                    void *__restrict__ out, long long num) {
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < num; o += (long long)gridDim.x * blockDim.x) {
        long long start; int row;
        map_output(m, o, &start, &row);
        const float *c = taps + (long long)row * m.row_stride;
        if (CPLX) {
            const float2 *pa = (const float2 *)a, *pb = (const float2 *)b;
            float re = 0.0f, im = 0.0f;
            if (start + T <= na) {   // common case: whole window in the first segment
                const float2 *x = pa + start;
                for (int k = 0; k < T; k++) { float2 v = __ldg(x + k); float t = __ldg(c + k);
                                              re = fmaf(t, v.x, re); im = fmaf(t, v.y, im); }
            } else {
                for (int k = 0; k < T; k++) { float2 v = seg_at<float2>(pa, na, pb, nb, start + k); float t = __ldg(c + k);
                                              re = fmaf(t, v.x, re); im = fmaf(t, v.y, im); }
            }
            ((float2 *)out)[o] = make_float2(re, im);
        } else {
            const float *pa = (const float *)a, *pb = (const float *)b;
            float acc = 0.0f;
            if (start + T <= na) {
                const float *x = pa + start;
                for (int k = 0; k < T; k++) acc = fmaf(__ldg(c + k), __ldg(x + k), acc);
            } else {
                for (int k = 0; k < T; k++) acc = fmaf(__ldg(c + k), seg_at<float>(pa, na, pb, nb, start + k), acc);
            }
            ((float *)out)[o] = acc;
        }
    }
}

// Tiled form of the same computation: a CTA stages the input span of 128 consecutive outputs (and the tap rows) in
// shared memory with coalesced loads, then each thread runs the identical sequential FMA chain out of shared memory.
// One global round trip per tile instead of one per tap: this is what the small ragged-tail / halo-boundary launches
// and every shape without a tuned kernel run on.
constexpr int TILE_OUT = 128;

template <bool CPLX>
__global__ void __launch_bounds__(TILE_OUT) k_fir_tile(OutMap m, int T, int n_rows, const float *__restrict__ taps,
                                                       const void *__restrict__ a, long long na,
                                                       const void *__restrict__ b, long long nb,
                                                       void *__restrict__ out, long long num) {
    typedef typename std::conditional<CPLX, float2, float>::type E;
    extern __shared__ __align__(16) unsigned char tile_smem[];
    const int row_floats = (m.ng == 0) ? T : m.row_stride;
    float *ts = reinterpret_cast<float *>(tile_smem);
    E *xs = reinterpret_cast<E *>(tile_smem + (((size_t)n_rows * row_floats * 4 + 15) / 16) * 16);
    for (int i = threadIdx.x; i < n_rows * row_floats; i += TILE_OUT) ts[i] = __ldg(taps + i);
    const long long n_tiles = (num + TILE_OUT - 1) / TILE_OUT;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long o0 = tile * TILE_OUT;
        const int n_o = (int)((num - o0) < TILE_OUT ? (num - o0) : TILE_OUT);
        long long start0, start_last; int row;
        map_output(m, o0, &start0, &row);
        map_output(m, o0 + n_o - 1, &start_last, &row);
        const int span = (int)(start_last - start0) + T;
        __syncthreads();   // previous tile fully consumed (and taps visible on the first trip)
        for (int i = threadIdx.x; i < span; i += TILE_OUT) xs[i] = seg_at<E>((const E *)a, na, (const E *)b, nb, start0 + i);
        __syncthreads();
        if ((int)threadIdx.x < n_o) {
            long long start;
            map_output(m, o0 + threadIdx.x, &start, &row);
            const E *x = xs + (start - start0);
            const float *c = ts + row * row_floats;
            if (CPLX) {
                float re = 0.0f, im = 0.0f;
#pragma unroll 8
                for (int k = 0; k < T; k++) { float2 v = ((const float2 *)x)[k]; float t = c[k];
                                              re = fmaf(t, v.x, re); im = fmaf(t, v.y, im); }
                ((float2 *)out)[o0 + threadIdx.x] = make_float2(re, im);
            } else {
                float acc = 0.0f;
#pragma unroll 8
                for (int k = 0; k < T; k++) acc = fmaf(c[k], ((const float *)x)[k], acc);
                ((float *)out)[o0 + threadIdx.x] = acc;
            }
        }
    }
}

static int launch_direct(Ctx *c, bool cplx, OutMap m, int T, const float *d_taps, Seg2 seg, void *d_out, long long num) {
    if (num <= 0) return SDR_OK;
    SDR_TRY(c->bind());
    // shared-memory need of the tiled kernel: tap rows + the widest input span of one tile
    const int n_rows = (m.ng == 0) ? 1 : m.ng;
    const int row_floats = (m.ng == 0) ? T : m.row_stride;
    long long span;
    if (m.ng == 0) span = (long long)(TILE_OUT - 1) * m.D + T;
    else           span = ((long long)(TILE_OUT - 1) / m.ng + 2) * m.sum_inc + T;
    size_t smem = (((size_t)n_rows * row_floats * 4 + 15) / 16) * 16 + (size_t)span * (cplx ? 8 : 4);
    static const bool no_tile = getenv("SDR_B200_NOTILE") != nullptr;   // debugging aid: force the direct kernel
    if (smem <= 200 * 1024 && !no_tile) {
        SDR_TRY(ring_attr(c, reinterpret_cast<const void *>(k_fir_tile<true>), 200 * 1024));
        SDR_TRY(ring_attr(c, reinterpret_cast<const void *>(k_fir_tile<false>), 200 * 1024));
        long long tiles = (num + TILE_OUT - 1) / TILE_OUT;
        long long cap = (long long)c->sm_count * (smem > 48 * 1024 ? 1 : 8);
        int grid = (int)(tiles < cap ? tiles : cap);
        if (cplx) k_fir_tile<true><<<grid, TILE_OUT, smem, c->s()>>>(m, T, n_rows, d_taps, seg.a, seg.na, seg.b, seg.nb, d_out, num);
        else      k_fir_tile<false><<<grid, TILE_OUT, smem, c->s()>>>(m, T, n_rows, d_taps, seg.a, seg.na, seg.b, seg.nb, d_out, num);
        SDR_LAUNCH_CHECK(c);
        return SDR_OK;
    }
    int grid = grid_for(num, 256, c->sm_count, 32);
    if (cplx) k_fir_direct<true><<<grid, 256, 0, c->s()>>>(m, T, d_taps, seg.a, seg.na, seg.b, seg.nb, d_out, num);
    else      k_fir_direct<false><<<grid, 256, 0, c->s()>>>(m, T, d_taps, seg.a, seg.na, seg.b, seg.nb, d_out, num);
    SDR_LAUNCH_CHECK(c);
    return SDR_OK;
}

int launch_fir_generic(Ctx *c, bool cplx, int T, int D, const float *d_taps, Seg2 seg, void *d_out, long long num) {
    OutMap m = {D, 0, 0, 0, nullptr, 0};
    return launch_direct(c, cplx, m, T, d_taps, seg, d_out, num);
}

int launch_resample_groups(Ctx *c, bool cplx, int taps_per_group, int row_stride, int g0, int ng, const int *d_prefix,
                           int sum_inc, const float *d_table, Seg2 seg, void *d_out, long long num) {
    OutMap m = {0, g0, ng, sum_inc, d_prefix, row_stride};
    return launch_direct(c, cplx, m, taps_per_group, d_table, seg, d_out, num);
}

// ---------------------------------------------------------------------------------------------------------------
// EXACT arithmetic: reproduces the reference's float summation order, unfused (the reference is built without
// -mfma, sdr.cabal:114): W independent lane accumulators exactly as the SIMD register holds them
// (common.h:43-72 real, :107-155 complex "2" form, :160-266 symmetric) then the reference's horizontal-add tree
// (common.h:12-29, :77-90).  Bit-identical to the scalar / SSE / AVX C variants; verification mode.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mac_unfused(float acc, float c, float x) { return __fadd_rn(acc, __fmul_rn(c, x)); }

__device__ __forceinline__ float tree(const float *l, int n) {
    // n = 1: l0 ; 2: l0+l1 ; 4: (l0+l1)+(l2+l3) ; 8: ((l0+l1)+(l2+l3)) + ((l4+l5)+(l6+l7))
    if (n == 1) return l[0];
    if (n == 2) return __fadd_rn(l[0], l[1]);
    if (n == 4) return __fadd_rn(__fadd_rn(l[0], l[1]), __fadd_rn(l[2], l[3]));
    return __fadd_rn(__fadd_rn(__fadd_rn(l[0], l[1]), __fadd_rn(l[2], l[3])),
                     __fadd_rn(__fadd_rn(l[4], l[5]), __fadd_rn(l[6], l[7])));
}

// layout: 0 = real W lanes (dotprod_R / sym_dotprod_R); 1 = complex, duplicated-coefficient form (dotprod_R over the
// interleaved floats: W/2 taps per step); 2 = complex "2" form (dotprod_C: accum1/accum2 halves); W in {1,4,8}.
// sym: T = HALF the taps, pre-add x[k] + x[2T-1-k].
template <bool CPLX>
__global__ void __launch_bounds__(128) k_fir_exact(OutMap m, int T, int W, int layout, int sym,
                                                   const float *__restrict__ taps,
                                                   const void *__restrict__ a, long long na,
                                                   const void *__restrict__ b, long long nb,
                                                   void *__restrict__ out, long long num) {
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < num; o += (long long)gridDim.x * blockDim.x) {
        long long start; int row;
        map_output(m, o, &start, &row);
        const float *c = taps + (long long)row * m.row_stride;
        float lr[8], li[8];
#pragma unroll
        for (int j = 0; j < 8; j++) lr[j] = li[j] = 0.0f;
        int nl = (!CPLX) ? W : (layout == 1 ? (W > 1 ? W / 2 : 1) : W);   // number of tap lanes
        for (int k = 0; k < T; k++) {
            float t = __ldg(c + k);
            int j = k % nl;
            if (CPLX) {
                float2 v = seg_at<float2>((const float2 *)a, na, (const float2 *)b, nb, start + k);
                if (sym) { float2 w = seg_at<float2>((const float2 *)a, na, (const float2 *)b, nb, start + 2 * T - 1 - k);
                           v.x = __fadd_rn(v.x, w.x); v.y = __fadd_rn(v.y, w.y); }
                // select lane with a switch-free unrolled compare so the arrays stay in registers
#pragma unroll
                for (int q = 0; q < 8; q++) if (q == j) { lr[q] = mac_unfused(lr[q], t, v.x); li[q] = mac_unfused(li[q], t, v.y); }
            } else {
                float v = seg_at<float>((const float *)a, na, (const float *)b, nb, start + k);
                if (sym) v = __fadd_rn(v, seg_at<float>((const float *)a, na, (const float *)b, nb, start + 2 * T - 1 - k));
#pragma unroll
                for (int q = 0; q < 8; q++) if (q == j) lr[q] = mac_unfused(lr[q], t, v);
            }
        }
        if (CPLX) {
            float re, im;
            if (layout == 2 && W > 1) {   // lanes = accum1 + accum2, then hadd_C over W/2 values
                int h = W / 2;
                float sr[4], si[4];
#pragma unroll
                for (int q = 0; q < 4; q++) { sr[q] = (q < h) ? __fadd_rn(lr[q], lr[q + h]) : 0.0f;
                                              si[q] = (q < h) ? __fadd_rn(li[q], li[q + h]) : 0.0f; }
                re = tree(sr, h); im = tree(si, h);
            } else {
                re = tree(lr, nl); im = tree(li, nl);
            }
            ((float2 *)out)[o] = make_float2(re, im);
        } else {
            ((float *)out)[o] = tree(lr, nl);
        }
    }
}

int launch_fir_exact(Ctx *c, bool cplx, OutMap m, int T, int W, int layout, int sym, const float *d_taps, Seg2 seg,
                     void *d_out, long long num) {
    if (num <= 0) return SDR_OK;
    if (!(W == 1 || W == 4 || W == 8)) return set_error(SDR_EINVAL, "exact arithmetic: lane width %d not in {1,4,8}", W);
    SDR_TRY(c->bind());
    int grid = grid_for(num, 128, c->sm_count, 32);
    if (cplx) k_fir_exact<true><<<grid, 128, 0, c->s()>>>(m, T, W, layout, sym, d_taps, seg.a, seg.na, seg.b, seg.nb, d_out, num);
    else      k_fir_exact<false><<<grid, 128, 0, c->s()>>>(m, T, W, layout, sym, d_taps, seg.a, seg.na, seg.b, seg.nb, d_out, num);
    SDR_LAUNCH_CHECK(c);
    return SDR_OK;
}

int launch_fir_exact_fir(Ctx *c, bool cplx, int T, int D, int W, int layout, int sym, const float *d_taps, Seg2 seg,
                         void *d_out, long long num) {
    OutMap m = {D, 0, 0, 0, nullptr, 0};
    return launch_fir_exact(c, cplx, m, T, W, layout, sym, d_taps, seg, d_out, num);
}

int launch_resample_exact(Ctx *c, bool cplx, int taps_per_group, int row_stride, int W, int layout, int g0, int ng,
                          const int *d_prefix, int sum_inc, const float *d_table, Seg2 seg, void *d_out, long long num) {
    OutMap m = {0, g0, ng, sum_inc, d_prefix, row_stride};
    return launch_fir_exact(c, cplx, m, taps_per_group, W, layout, 0, d_table, seg, d_out, num);
}

// ---------------------------------------------------------------------------------------------------------------
// element-wise kernels: 16-byte vector path when both pointers are 16-byte aligned, scalar otherwise
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cvt_u8(unsigned b) { return __fmul_rn(__fsub_rn((float)b, 128.0f), 1.0f / 128.0f); }

// one 32-bit word (4 bytes) in, one float4 out per thread and trip: both sides fully coalesced; 4 trips in flight
__global__ void __launch_bounds__(256) k_convert_u8_vec(const uint32_t *__restrict__ in, float4 *__restrict__ out, long long n4) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) w[q] = __ldg(in + i + q * stride);
#pragma unroll
        for (int q = 0; q < 4; q++)
            out[i + q * stride] = make_float4(cvt_u8(w[q] & 0xff), cvt_u8((w[q] >> 8) & 0xff), cvt_u8((w[q] >> 16) & 0xff), cvt_u8(w[q] >> 24));
    }
    for (; i < n4; i += stride) {
        uint32_t w = __ldg(in + i);
        out[i] = make_float4(cvt_u8(w & 0xff), cvt_u8((w >> 8) & 0xff), cvt_u8((w >> 16) & 0xff), cvt_u8(w >> 24));
    }
}
__global__ void __launch_bounds__(256) k_convert_u8(const uint8_t *__restrict__ in, float *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = cvt_u8(in[i]);
}

int launch_convert_u8(Ctx *c, const uint8_t *d_in, float *d_out, long long n) {
    if (n <= 0) return SDR_OK;
    SDR_TRY(c->bind());
    long long nv = 0;
    if ((((uintptr_t)d_in) & 3) == 0 && (((uintptr_t)d_out) & 15) == 0) nv = n / 4;
    if (nv) { k_convert_u8_vec<<<grid_for(nv, 256, c->sm_count, 8), 256, 0, c->s()>>>((const uint32_t *)d_in, (float4 *)d_out, nv);
              SDR_LAUNCH_CHECK(c); }
    long long rest = n - nv * 4;
    if (rest) { k_convert_u8<<<grid_for(rest, 256, c->sm_count), 256, 0, c->s()>>>(d_in + nv * 4, d_out + nv * 4, rest);
                SDR_LAUNCH_CHECK(c); }
    return SDR_OK;
}

__global__ void __launch_bounds__(256) k_convert_i16(const int16_t *__restrict__ in, float *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = __fmul_rn((float)in[i], 1.0f / 2048.0f);
}
int launch_convert_i16(Ctx *c, const int16_t *d_in, float *d_out, long long n) {
    if (n <= 0) return SDR_OK;
    SDR_TRY(c->bind());
    k_convert_i16<<<grid_for(n, 256, c->sm_count), 256, 0, c->s()>>>(d_in, d_out, n);
    SDR_LAUNCH_CHECK(c);
    return SDR_OK;
}

__global__ void __launch_bounds__(256) k_convert_tx(const float *__restrict__ in, int16_t *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float val = __fmul_rn(__fadd_rn(in[i], 1.0f), 2048.0f);
        // (int16_t)val on x86 = cvttss2si then truncation to 16 bits (convert.c:93); out-of-int32-range -> 0x80000000
        int wide = (val >= 2147483648.0f || val < -2147483648.0f || val != val) ? (int)0x80000000 : (int)val;
        int16_t res = (int16_t)wide;
        res = (int16_t)(res - 2048);
        if (res > 2047) res = 2047;
        if (res < -2048) res = -2048;
        out[i] = res;
    }
}
int launch_convert_tx(Ctx *c, const float *d_in, int16_t *d_out, long long n) {
    if (n <= 0) return SDR_OK;
    SDR_TRY(c->bind());
    k_convert_tx<<<grid_for(n, 256, c->sm_count), 256, 0, c->s()>>>(d_in, d_out, n);
    SDR_LAUNCH_CHECK(c);
    return SDR_OK;
}

__global__ void __launch_bounds__(256) k_scale_vec(float k, const float4 *__restrict__ in, float4 *__restrict__ out, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(in + i);
        out[i] = make_float4(__fmul_rn(v.x, k), __fmul_rn(v.y, k), __fmul_rn(v.z, k), __fmul_rn(v.w, k));
    }
}
__global__ void __launch_bounds__(256) k_scale(float k, const float *__restrict__ in, float *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = __fmul_rn(in[i], k);
}
int launch_scale(Ctx *c, float k, const float *d_in, float *d_out, long long n) {
    if (n <= 0) return SDR_OK;
    SDR_TRY(c->bind());
    long long nv = 0;
    if ((((uintptr_t)d_in | (uintptr_t)d_out) & 15) == 0) nv = n / 4;
    if (nv) { k_scale_vec<<<grid_for(nv, 256, c->sm_count), 256, 0, c->s()>>>(k, (const float4 *)d_in, (float4 *)d_out, nv);
              SDR_LAUNCH_CHECK(c); }
    long long rest = n - nv * 4;
    if (rest) { k_scale<<<grid_for(rest, 256, c->sm_count), 256, 0, c->s()>>>(k, d_in + nv * 4, d_out + nv * 4, rest);
                SDR_LAUNCH_CHECK(c); }
    return SDR_OK;
}

// 4 consecutive samples per thread and trip: two 16-byte loads (+ one L1-resident re-read for the predecessor of the
// first sample), four independent discriminators, one 16-byte store.  Two trips are in flight per thread (all six loads
// are issued before the first discriminator): with one, ncu showed the kernel waiting on memory (long_scoreboard 12.8
// warps per issue at 75 % of DRAM bandwidth).
__device__ __forceinline__ float4 fm_demod_group(const float4 a, const float4 b, const float2 l) {
    const float2 s0 = make_float2(a.x, a.y), s1 = make_float2(a.z, a.w), s2 = make_float2(b.x, b.y), s3 = make_float2(b.z, b.w);
    return make_float4(fm_phase(s0, l), fm_phase(s1, s0), fm_phase(s2, s1), fm_phase(s3, s2));
}
__global__ void __launch_bounds__(256) k_fm_demod4(float last_re, float last_im, const float2 *__restrict__ last_ptr,
                                                   const float4 *__restrict__ in, float4 *__restrict__ out, long long n4) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
        const long long i2 = i + stride;
        const bool      two = i2 < n4;
        const float4 a = __ldg(in + 2 * i), b = __ldg(in + 2 * i + 1);
        float4 a2 = a, b2 = b, p2 = a;
        if (two) { a2 = __ldg(in + 2 * i2); b2 = __ldg(in + 2 * i2 + 1); p2 = __ldg(in + 2 * i2 - 1); }
        float2 l;
        if (i == 0) l = last_ptr ? *last_ptr : make_float2(last_re, last_im);
        else { const float4 p = __ldg(in + 2 * i - 1); l = make_float2(p.z, p.w); }
        out[i] = fm_demod_group(a, b, l);
        if (two) out[i2] = fm_demod_group(a2, b2, make_float2(p2.z, p2.w));
    }
}
__global__ void __launch_bounds__(256) k_fm_demod(float last_re, float last_im, const float2 *__restrict__ last_ptr,
                                                  const float2 *__restrict__ in, float *__restrict__ out, long long first, long long n) {
    for (long long i = first + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float2 s = __ldg(in + i);
        float2 l = (i == 0) ? (last_ptr ? *last_ptr : make_float2(last_re, last_im)) : __ldg(in + i - 1);
        out[i] = fm_phase(s, l);
    }
}
static int launch_fm_demod_any(Ctx *c, float last_re, float last_im, const float *d_last, const float *d_in, float *d_out, long long n) {
    if (n <= 0) return SDR_OK;
    SDR_TRY(c->bind());
    long long n4 = 0;
    if ((((uintptr_t)d_in | (uintptr_t)d_out) & 15) == 0) n4 = n / 4;
    if (n4) { k_fm_demod4<<<grid_for(n4, 256, c->sm_count, 8), 256, 0, c->s()>>>(last_re, last_im, (const float2 *)d_last, (const float4 *)d_in,
                                                                               (float4 *)d_out, n4);
              SDR_LAUNCH_CHECK(c); }
    if (4 * n4 < n) { k_fm_demod<<<grid_for(n - 4 * n4, 256, c->sm_count), 256, 0, c->s()>>>(last_re, last_im, (const float2 *)d_last,
                                                                                            (const float2 *)d_in, d_out, 4 * n4, n);
                      SDR_LAUNCH_CHECK(c); }
    return SDR_OK;
}
int launch_fm_demod(Ctx *c, float last_re, float last_im, const float *d_in, float *d_out, long long n) {
    return launch_fm_demod_any(c, last_re, last_im, nullptr, d_in, d_out, n);
}
// streaming form: the previous buffer's final sample is read from device memory (fmDemod's carried state, Demod.hs:46)
int launch_fm_demod_carry(Ctx *c, const float *d_last, const float *d_in, float *d_out, long long n) {
    return launch_fm_demod_any(c, 0.0f, 0.0f, d_last, d_in, d_out, n);
}

// dcBlocker / dcBlockingFilter: kernels_dc.cu

// ---------------------------------------------------------------------------------------------------------------
// synthetic streams (counter-based, keyed on the GLOBAL element index) and measurement helpers
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {   // "lowbias32" integer finaliser
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ float noise_at(uint64_t idx, uint32_t seed) {
    uint32_t h1 = mix32((uint32_t)idx ^ mix32((uint32_t)(idx >> 32) ^ seed));
    uint32_t h2 = mix32(h1 ^ 0x9E3779B9U);
    int s = (int)((h1 & 0xffff) + (h1 >> 16) + (h2 & 0xffff) + (h2 >> 16)) - 131070;   // Irwin-Hall(4), zero mean
    return (float)s * 2.6429e-5f;                                                      // ~unit variance; one rounding
}
__global__ void __launch_bounds__(256) k_synth_noise(float *__restrict__ out, long long n, long long first, uint32_t seed) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = noise_at((uint64_t)(first + i), seed);
}
int launch_synth_noise(Ctx *c, float *d_out, long long n, long long first, uint32_t seed) {
    if (n <= 0) return SDR_OK;
    SDR_TRY(c->bind());
    k_synth_noise<<<grid_for(n, 256, c->sm_count), 256, 0, c->s()>>>(d_out, n, first, seed);
    SDR_LAUNCH_CHECK(c);
    return SDR_OK;
}
__global__ void __launch_bounds__(256) k_synth_bytes(uint8_t *__restrict__ out, long long n, long long first, uint32_t seed) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint64_t g = (uint64_t)(first + i), w = g >> 2;
        uint32_t h = mix32((uint32_t)w ^ mix32((uint32_t)(w >> 32) ^ seed ^ 0xB5297A4DU));
        out[i] = (uint8_t)(h >> (8 * (g & 3)));
    }
}
int launch_synth_bytes(Ctx *c, uint8_t *d_out, long long n, long long first, uint32_t seed) {
    if (n <= 0) return SDR_OK;
    SDR_TRY(c->bind());
    k_synth_bytes<<<grid_for(n, 256, c->sm_count), 256, 0, c->s()>>>(d_out, n, first, seed);
    SDR_LAUNCH_CHECK(c);
    return SDR_OK;
}

// position-weighted wrap-around checksum: sum_i (w[i] + 1) * (2 * (first + i) + 1)  mod 2^64
__global__ void __launch_bounds__(256) k_checksum32(const uint32_t *__restrict__ buf, long long n, long long first,
                                                    unsigned long long *sum) {
    unsigned long long s = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s += ((unsigned long long)buf[i] + 1ULL) * (2ULL * (unsigned long long)(first + i) + 1ULL);
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum, s);
}
int launch_checksum32(Ctx *c, const uint32_t *d_buf, long long n, long long first, unsigned long long *d_sum) {
    SDR_TRY(c->bind());
    SDR_CUDA(cudaMemsetAsync(d_sum, 0, 8, c->s()));
    if (n <= 0) return SDR_OK;
    k_checksum32<<<grid_for(n, 256, c->sm_count), 256, 0, c->s()>>>(d_buf, n, first, d_sum);
    SDR_LAUNCH_CHECK(c);
    return SDR_OK;
}

__global__ void __launch_bounds__(256) k_fill(uint4 *__restrict__ p, long long n16, unsigned v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        p[i] = make_uint4(v, v, v, v);
}
int launch_fill(Ctx *c, void *d, size_t bytes) {
    SDR_TRY(c->bind());
    k_fill<<<grid_for((long long)(bytes / 16), 256, c->sm_count), 256, 0, c->s()>>>((uint4 *)d, (long long)(bytes / 16), 0x5d2b200u);
    SDR_LAUNCH_CHECK(c);
    return SDR_OK;
}

}  // namespace sdr
