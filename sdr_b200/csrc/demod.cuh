// demod.cuh -- the FM discriminator of one sample pair, shared by the stand-alone kernel (kernels_generic.cu) and the
// fused front end (kernels_fm.cu) so both produce identical bits.  Reference: hs_sources/SDR/Demod.hs:21-36.
#pragma once
#include "common.cuh"

namespace sdr {

// GHC's class-default atan2 for Float (GHC.Float) is a case analysis around atan(y/x) (oracle/sdr_oracle.c hs_atan2f).
// Evaluated here without the IEEE division and libm atanf it is written with (~45 of the discriminator's ~55
// instructions): t = min(|x|,|y|) / max(|x|,|y|) in [0,1] by the approximate reciprocal, atan(t) = t*P(t^2) with a
// degree-7 minimax P (3.7e-8 absolute), then the octant is unfolded with the same constants the reference adds
// (pi/2 - p, pi - p, sign of y).  Same value table as the reference for signed zeros, infinities and NaN (NaN-propagating
// min/max); measured against the oracle over noise at six scales, the u8 grid and an edge-value soup: at most 1 ulp of
// pi (2.4e-7) apart -- the 1e-5 bar with 40x margin.  Selects only, no branches: a lane of the fused front end
// evaluates eight of these back to back.  Explicit intrinsics so that every translation unit compiles the same chain.
__device__ __forceinline__ float hs_atan2f_dev(float y, float x) {
    const float pi = 3.14159265358979323846f, half_pi = 1.57079632679489661923f;
    const float ax = fabsf(x), ay = fabsf(y);
    float mx, mn;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(mx) : "f"(ax), "f"(ay));
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(mn) : "f"(ax), "f"(ay));
    // __fdividef rescales denormal divisors itself but returns 0 for divisors above 2^126: bring huge products down
    const float sc = (mx > 0x1p100f) ? 0x1p-64f : 1.0f;
    const float t = __fdividef(__fmul_rn(mn, sc), __fmul_rn(mx, sc));
    const float u = __fmul_rn(t, t);
    float p = -0.00405447837262076f;
    p = __fmaf_rn(p, u, 0.021862652422097704f);
    p = __fmaf_rn(p, u, -0.055911910273229344f);
    p = __fmaf_rn(p, u, 0.09642168717203611f);
    p = __fmaf_rn(p, u, -0.13908619173401218f);
    p = __fmaf_rn(p, u, 0.19946563758352057f);
    p = __fmaf_rn(p, u, -0.33329860637118436f);
    p = __fmaf_rn(p, u, 0.9999993355476848f);
    p = __fmul_rn(p, t);
    p = (ay > ax) ? __fsub_rn(half_pi, p) : p;
    p = (x < 0.0f) ? __fsub_rn(pi, p) : p;
    return copysignf(p, y);
}

// phase(s * conj(last)) exactly as Haskell evaluates it: (a:+b)*(c:+d) = (a*c - b*d) :+ (a*d + b*c) with
// conj(last) = lr :+ (-li), every product rounded separately (no FMA contraction); phase (0:+0) = 0 (Data.Complex)
__device__ __forceinline__ float fm_phase(float2 s, float2 l) {
    const float nli = -l.y;
    const float re = __fsub_rn(__fmul_rn(s.x, l.x), __fmul_rn(s.y, nli));
    const float im = __fadd_rn(__fmul_rn(s.x, nli), __fmul_rn(s.y, l.x));
    return (re == 0.0f && im == 0.0f) ? 0.0f : hs_atan2f_dev(im, re);
}

}  // namespace sdr
