// demod.cuh -- the FM discriminator of one sample pair, shared by the stand-alone kernel (kernels_generic.cu) and the
// fused front end (kernels_fm.cu) so both produce identical bits.  Reference: hs_sources/SDR/Demod.hs:21-36.
#pragma once
#include "common.cuh"

namespace sdr {

// GHC's class-default atan2 for Float (GHC.Float), built on atan: see oracle/sdr_oracle.c hs_atan2f
__device__ __forceinline__ bool neg_zero(float v) { return v == 0.0f && signbit(v); }
// Branch-free (selects only): a lane evaluates several of these back to back in the fused front end and divergent
// control flow would serialise them; the value table is exactly the reference's case analysis.
__device__ __forceinline__ float hs_atan2f_dev(float y, float x) {
    const float pi = 3.14159265358979323846f;
    const bool flip = (x <= 0 && y < 0) || (x < 0 && neg_zero(y)) || (neg_zero(x) && neg_zero(y));
    const float yf = flip ? -y : y;   // the reference recurses once with -y and negates the result
    const float a = atanf(__fdiv_rn(yf, x));
    float r = x + yf;                                              // otherwise (NaN operands)
    r = (x == 0 && yf == 0) ? yf : r;
    r = (yf == 0 && (x < 0 || neg_zero(x))) ? pi : r;
    r = (x < 0 && yf > 0) ? __fadd_rn(pi, a) : r;
    r = (x == 0 && yf > 0) ? (pi / 2) : r;
    r = (x > 0) ? a : r;
    return flip ? -r : r;
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
