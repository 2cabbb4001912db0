// fir_ffa.cuh -- 2-parallel fast-FIR evaluation of R consecutive outputs of a real stride-1 FIR (reference
// filterAVXRR / filterAVXSymmetricRR, c_sources/filter.c:37-68) from one lane's window.
//
// The direct form spends T multiply-adds per output and the 64-tap filter is bound by the FP32 pipe, not by HBM
// (DESIGN.md section 4.2).  Splitting taps and samples into even and odd phases (h0[j] = h[2j], h1[j] = h[2j+1]),
//     A[m] = sum_j h0[j]       x[2m + 2j]                       m = 0 .. R/2
//     B[m] = sum_j h1[j]       x[2m + 2j + 1]                   m = 0 .. R/2 - 1
//     C[m] = sum_j (h0+h1)[j] (x[2m + 2j + 1] + x[2m + 2j + 2])
//     y[2m]     = A[m] + B[m]
//     y[2m + 1] = C[m] - A[m + 1] - B[m]            (C[m] = H0*X1[m] + A[m+1] + B[m] + H1*X0[m+1])
// needs three half-length sub-filters per two outputs instead of four: (3 R/2 + 1) T/2 multiply-adds plus about
// (R + T)/2 + 3 R/2 additions for R outputs -- 1063 FP32-pipe operations instead of 1280 at T = 64, R = 20.  The sums
// are associated differently from the reference's, so results agree with it to rounding, not bit for bit: measured
// 1.5e-6 of the output scale against 7e-7 for the direct form (bar: 1e-5; tests/test_fir_ffa_model.py).
//
// __host__ __device__ so that the CPU test runs the very index arithmetic the kernel is unrolled from.
#pragma once
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define SDR_FFA_HD __host__ __device__ __forceinline__
#else
#define SDR_FFA_HD inline
#endif

namespace sdr {

SDR_FFA_HD float ffa_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}

// w: the lane's window as WIN4 = (R + T - 1 + 3) / 4 float4 chunks starting at the position of output 0;
// h0, h1, hs: the T/2 even taps, odd taps and their sums; y: the R outputs.  All loops have compile-time bounds and
// indices, so on the device everything lives in registers.
template <int T, int R>
SDR_FFA_HD void fir_ffa_lane(const float4 *w, const float (&h0)[T / 2], const float (&h1)[T / 2], const float (&hs)[T / 2],
                             float (&y)[R]) {
    static_assert(T % 2 == 0 && R % 2 == 0, "even tap count and an even number of outputs per lane");
    constexpr int H = T / 2, P = R / 2, WIN4 = (R + T - 1 + 3) / 4;
    float A[P + 1], B[P], C[P];
#pragma unroll
    for (int m = 0; m <= P; m++) A[m] = 0.0f;
#pragma unroll
    for (int m = 0; m < P; m++) { B[m] = 0.0f; C[m] = 0.0f; }
    float prev_w = 0.0f;
#pragma unroll
    for (int c4 = 0; c4 < WIN4; c4++) {
        const float4 v = w[c4];
        const float  e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int idx = 4 * c4 + i;
            if (idx % 2 == 0) {
                const int p = idx / 2;
#pragma unroll
                for (int m = 0; m <= P; m++) {
                    const int j = p - m;
                    if (j >= 0 && j < H) A[m] = ffa_fma(h0[j], e[i], A[m]);
                }
                if (idx >= 2) {
                    const float s = (i == 0 ? prev_w : e[i == 0 ? 0 : i - 1]) + e[i];   // x[idx - 1] + x[idx]
                    const int   q = p - 1;
#pragma unroll
                    for (int m = 0; m < P; m++) {
                        const int j = q - m;
                        if (j >= 0 && j < H) C[m] = ffa_fma(hs[j], s, C[m]);
                    }
                }
            } else {
                const int p = (idx - 1) / 2;
#pragma unroll
                for (int m = 0; m < P; m++) {
                    const int j = p - m;
                    if (j >= 0 && j < H) B[m] = ffa_fma(h1[j], e[i], B[m]);
                }
            }
        }
        prev_w = e[3];
    }
#pragma unroll
    for (int m = 0; m < P; m++) {
        y[2 * m] = A[m] + B[m];
        y[2 * m + 1] = (C[m] - A[m + 1]) - B[m];
    }
}

}  // namespace sdr
