/*
 * sdr_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the arithmetic of the reference's native hot path
 * (adamwalker/sdr c_sources/{common.h,filter.c,decimate.c,resample.c,convert.c,scale.c}) plus the pure-Haskell
 * pieces of the same path that have no C (cross-buffer kernels, fmDemod).  Plain C, no intrinsics: every SIMD
 * variant of the reference is restated as "W independent lane accumulators + the reference's horizontal-add
 * tree", which reproduces the reference's float summation ORDER and therefore its bits.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object.  The product (sdr_b200/) never links, imports or calls it.
 *
 * Parity pin: tests/test_oracle_golden.py checks every function here bit-for-bit against golden vectors produced
 * by the UNMODIFIED reference C compiled with its own flags (oracle/Makefile -> oracle/_ref/libsdrref.so;
 * generator tests/golden/make_golden.py).  Functions restating Haskell code (o_*Cross*, o_fmDemod*) have no
 * reference executable here (no GHC in this image): "parity unpinned" for those -- they follow the cited lines.
 *
 * Build with -ffp-contract=off: the reference is built without -mfma (sdr.cabal:114), i.e. mul and add round
 * separately everywhere (common.h:53,69,122-123,150-151,175,197).
 */

#include <stdint.h>
#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------
 * Horizontal adds (common.h:12-29 real, :77-90 complex)
 * ---------------------------------------------------------------------------------------------------------- */

/* sse_hadd_R common.h:12-16: hadd twice => (l0+l1)+(l2+l3) */
static float hadd_R4(const float *l) { return (l[0] + l[1]) + (l[2] + l[3]); }

/* avx_hadd_R common.h:18-29: each 128-bit half reduced as above, then low + high */
static float hadd_R8(const float *l) { return hadd_R4(l) + hadd_R4(l + 4); }

/* sse_hadd_C common.h:77-80: shuffle to (l0,l2,l1,l3) then hadd => re=l0+l2, im=l1+l3 */
static void hadd_C4(const float *l, float *out) {
    out[0] = l[0] + l[2];
    out[1] = l[1] + l[3];
}

/* avx_hadd_C common.h:82-90: re=(l0+l2)+(l4+l6), im=(l1+l3)+(l5+l7) */
static void hadd_C8(const float *l, float *out) {
    out[0] = (l[0] + l[2]) + (l[4] + l[6]);
    out[1] = (l[1] + l[3]) + (l[5] + l[7]);
}

/* ------------------------------------------------------------------------------------------------------------
 * Dot products.  `lanes` receives the W lane accumulators exactly as the SIMD register would hold them.
 * The SIMD loops step W elements with no tail handling (common.h:47,62): `num` is rounded up to W here the same
 * way the hardware loop would run, so callers must supply padded coefficient arrays just as the reference's
 * Haskell side does (Filter.hs:169,284,324).
 * ---------------------------------------------------------------------------------------------------------- */

/* dotprod_R common.h:34-41 */
static float dot_R1(int num, const float *a, const float *b) {
    float accum = 0;
    for (int i = 0; i < num; i++) accum += a[i] * b[i];
    return accum;
}

/* sse_dotprod_R common.h:43-56 (W=4), avx_dotprod_R common.h:58-72 (W=8) */
static void dot_RW(int W, int num, const float *a, const float *b, float *lanes) {
    for (int j = 0; j < W; j++) lanes[j] = 0.0f;
    for (int i = 0; i < num; i += W)
        for (int j = 0; j < W; j++) lanes[j] = lanes[j] + a[i + j] * b[i + j];
}

/* dotprod_C common.h:95-105: a = coeffs (num of them), b = interleaved complex data */
static void dot_C1(int num, const float *a, const float *b, float *result) {
    float real = 0, imag = 0;
    for (int i = 0; i < num; i++) {
        real += b[2 * i] * a[i];
        imag += b[2 * i + 1] * a[i];
    }
    result[0] = real;
    result[1] = imag;
}

/* sse_dotprod_C common.h:107-127 (W=4) and avx_dotprod_C common.h:129-155 (W=8).
 * Per iteration W coefficients; accum1 takes the first W/2 taps of the group (each tap on a (re,im) lane pair),
 * accum2 the second W/2; the two accumulators are added lane-wise at the end. */
static void dot_CW(int W, int num, const float *coeffs, const float *x, float *lanes) {
    float a1[8], a2[8];
    int h = W / 2;
    for (int j = 0; j < W; j++) a1[j] = a2[j] = 0.0f;
    for (int i = 0; i < num; i += W) {
        for (int t = 0; t < h; t++) {
            float c1 = coeffs[i + t], c2 = coeffs[i + h + t];
            a1[2 * t]     = a1[2 * t]     + c1 * x[2 * (i + t)];
            a1[2 * t + 1] = a1[2 * t + 1] + c1 * x[2 * (i + t) + 1];
            a2[2 * t]     = a2[2 * t]     + c2 * x[2 * (i + h + t)];
            a2[2 * t + 1] = a2[2 * t + 1] + c2 * x[2 * (i + h + t) + 1];
        }
    }
    for (int j = 0; j < W; j++) lanes[j] = a1[j] + a2[j];
}

/* sse_sym_dotprod_R common.h:160-179 (W=4), avx_sym_dotprod_R common.h:181-201 (W=8).
 * num = HALF the tap count; lane j of iteration i: c[i+j] * (x[i+j] + x[2num-1-i-j]). */
static void dot_symRW(int W, int num, const float *a, const float *b, float *lanes) {
    for (int j = 0; j < W; j++) lanes[j] = 0.0f;
    for (int i = 0; i < num; i += W)
        for (int j = 0; j < W; j++) lanes[j] = lanes[j] + a[i + j] * (b[i + j] + b[2 * num - 1 - i - j]);
}

/* sse_sym_dotprod_C common.h:206-232 (W=4), avx_sym_dotprod_C common.h:235-266 (W=8).
 * num = HALF the tap count, x interleaved complex.  Tap t=i+q pairs sample t with sample 2num-1-t. */
static void dot_symCW(int W, int num, const float *coeffs, const float *x, float *lanes) {
    float a1[8], a2[8];
    int h = W / 2;
    for (int j = 0; j < W; j++) a1[j] = a2[j] = 0.0f;
    for (int i = 0; i < num; i += W) {
        for (int t = 0; t < h; t++) {
            int k1 = i + t, k2 = i + h + t;
            int m1 = 2 * num - 1 - k1, m2 = 2 * num - 1 - k2;
            float c1 = coeffs[k1], c2 = coeffs[k2];
            a1[2 * t]     = a1[2 * t]     + c1 * (x[2 * k1]     + x[2 * m1]);
            a1[2 * t + 1] = a1[2 * t + 1] + c1 * (x[2 * k1 + 1] + x[2 * m1 + 1]);
            a2[2 * t]     = a2[2 * t]     + c2 * (x[2 * k2]     + x[2 * m2]);
            a2[2 * t + 1] = a2[2 * t + 1] + c2 * (x[2 * k2 + 1] + x[2 * m2 + 1]);
        }
    }
    for (int j = 0; j < W; j++) lanes[j] = a1[j] + a2[j];
}

/* ------------------------------------------------------------------------------------------------------------
 * One output of each reference kernel shape.  `variant` selects the reference function family member.
 * ---------------------------------------------------------------------------------------------------------- */

enum {
    V_SCALAR = 0, /* filterRR / filterRC / decimateRR / decimateRC / resample2RR / resample2RC          */
    V_SSE    = 1, /* *SSERR, *SSERC (dup coeffs)                                                        */
    V_AVX    = 2, /* *AVXRR, *AVXRC (dup coeffs)                                                        */
    V_SSE2   = 3, /* *SSERC2, resampleSSERC (sse_dotprod_C, plain coeffs)                               */
    V_AVX2   = 4, /* *AVXRC2, resampleAVXRC (avx_dotprod_C, plain coeffs)                               */
    V_SSESYM = 5, /* *SSESymmetricRR / RC (half coeffs)                                                 */
    V_AVXSYM = 6  /* *AVXSymmetricRR / RC (half coeffs)                                                 */
};

/* real data; numCoeffs as the reference function receives it (half length for the symmetric variants) */
static float one_R(int variant, int numCoeffs, const float *coeffs, const float *start) {
    float l[8];
    switch (variant) {
    case V_SCALAR: return dot_R1(numCoeffs, coeffs, start);                       /* filter.c:16-22   */
    case V_SSE:    dot_RW(4, numCoeffs, coeffs, start, l); return hadd_R4(l);     /* filter.c:27-35   */
    case V_AVX:    dot_RW(8, numCoeffs, coeffs, start, l); return hadd_R8(l);     /* filter.c:37-45   */
    case V_SSESYM: dot_symRW(4, numCoeffs, coeffs, start, l); return hadd_R4(l);  /* filter.c:50-58   */
    case V_AVXSYM: dot_symRW(8, numCoeffs, coeffs, start, l); return hadd_R8(l);  /* filter.c:60-68   */
    }
    return NAN;
}

/* complex data; start points at interleaved floats.  For V_SSE/V_AVX numCoeffs is the DUPLICATED length (2T) and
 * coeffs holds each tap twice (Filter.hs:206,326; FilterInternal.hs:177). */
static void one_C(int variant, int numCoeffs, const float *coeffs, const float *start, float *out) {
    float l[8];
    switch (variant) {
    case V_SCALAR: dot_C1(numCoeffs, coeffs, start, out); return;                     /* filter.c:74-80   */
    case V_SSE:    dot_RW(4, numCoeffs, coeffs, start, l); hadd_C4(l, out); return;   /* filter.c:86-94   */
    case V_AVX:    dot_RW(8, numCoeffs, coeffs, start, l); hadd_C8(l, out); return;   /* filter.c:106-114 */
    case V_SSE2:   dot_CW(4, numCoeffs, coeffs, start, l); hadd_C4(l, out); return;   /* filter.c:96-104  */
    case V_AVX2:   dot_CW(8, numCoeffs, coeffs, start, l); hadd_C8(l, out); return;   /* filter.c:116-124 */
    case V_SSESYM: dot_symCW(4, numCoeffs, coeffs, start, l); hadd_C4(l, out); return;/* filter.c:129-137 */
    case V_AVXSYM: dot_symCW(8, numCoeffs, coeffs, start, l); hadd_C8(l, out); return;/* filter.c:139-147 */
    }
    out[0] = out[1] = NAN;
}

/* ------------------------------------------------------------------------------------------------------------
 * Filters and decimators.  filter == decimate with factor 1 (filter.c vs decimate.c differ only in the stride).
 * ---------------------------------------------------------------------------------------------------------- */

/* decimate.c:16-68 (RR family); filter.c:16-68 with factor = 1 */
void o_decimateR(int variant, int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf,
                 float *outBuf) {
    for (int i = 0, k = 0; i < num; i++, k += factor) outBuf[i] = one_R(variant, numCoeffs, coeffs, inBuf + k);
}

/* decimate.c:73-146 (RC family); filter.c:74-147 with factor = 1.  k advances 2*factor floats per output. */
void o_decimateC(int variant, int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf,
                 float *outBuf) {
    for (int i = 0, k = 0; i < num * 2; i += 2, k += factor * 2)
        one_C(variant, numCoeffs, coeffs, inBuf + k, outBuf + i);
}

void o_filterR(int variant, int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    o_decimateR(variant, num, 1, numCoeffs, coeffs, inBuf, outBuf);
}

void o_filterC(int variant, int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    o_decimateC(variant, num, 1, numCoeffs, coeffs, inBuf, outBuf);
}

/* dcBlocker filter.c:152-161.  0.997 is a double literal: the product and sum are evaluated in double and
 * rounded to float on assignment. */
void o_dcBlocker(int num, float lastSample, float lastOutput, float *finalSample, float *finalOutput,
                 const float *inBuf, float *outBuf) {
    for (int i = 0; i < num; i++) {
        lastOutput = (float)((double)(inBuf[i] - lastSample) + 0.997 * (double)lastOutput);
        outBuf[i]  = lastOutput;
        lastSample = inBuf[i];
    }
    *finalSample = lastSample;
    *finalOutput = lastOutput;
}

/* ------------------------------------------------------------------------------------------------------------
 * Resamplers
 * ---------------------------------------------------------------------------------------------------------- */

/* resampleRR resample.c:16-32 (legacy single-array form) */
void o_resampleRR(int buf_size, int coeff_size, int interpolation, int decimation, int filter_offset,
                  const float *coeffs, const float *in_buf, float *out_buf) {
    int input_offset = 0;
    for (int k = 0; k < buf_size; k++) {
        float accum = 0;
        for (int l = 0, j = filter_offset; j < coeff_size; l++, j += interpolation)
            accum += in_buf[input_offset + l] * coeffs[j];
        int filter_offset_new = interpolation - 1 - (decimation - filter_offset - 1) % interpolation;
        input_offset += (decimation - filter_offset - 1) / interpolation + 1;
        filter_offset = filter_offset_new;
        out_buf[k]    = accum;
    }
}

/* resample2RR / resampleSSERR / resampleAVXRR resample.c:34-87.  coeffs is the flattened group table:
 * group g occupies coeffs[g*group_stride .. +num_coeffs). Returns the next group. */
int o_resampleR(int variant, int buf_size, int num_coeffs, int starting_group, int num_groups,
                const int *increments, const float *coeffs, int group_stride, const float *in_buf,
                float *out_buf) {
    int          group = starting_group;
    const float *start = in_buf;
    for (int i = 0; i < buf_size; i++) {
        out_buf[i] = one_R(variant, num_coeffs, coeffs + (long)group * group_stride, start);
        start += increments[group];
        group++;
        if (group == num_groups) group = 0;
    }
    return group;
}

/* resample2RC / resampleSSERC / resampleAVXRC resample.c:89-142: the SIMD complex resamplers use the
 * sse/avx_dotprod_C ("2") form with plain (un-duplicated) coefficients => variants V_SCALAR, V_SSE2, V_AVX2. */
int o_resampleC(int variant, int buf_size, int num_coeffs, int starting_group, int num_groups,
                const int *increments, const float *coeffs, int group_stride, const float *in_buf,
                float *out_buf) {
    int          group = starting_group;
    const float *start = in_buf;
    for (int i = 0; i < buf_size * 2; i += 2) {
        one_C(variant, num_coeffs, coeffs + (long)group * group_stride, start, out_buf + i);
        start += 2 * increments[group];
        group++;
        if (group == num_groups) group = 0;
    }
    return group;
}

/* ------------------------------------------------------------------------------------------------------------
 * Cross-buffer kernels (pure Haskell in the reference): strict left-to-right single-accumulator sums over
 * `drop i lastBuf ++ nextBuf` zipped with the coefficient vector (zipWith stops at the shorter list).
 * VG.sum = foldl' (+) 0: ((0 + p0) + p1) + ...   Mult (Complex a) a multiplies re and im by the tap separately
 * (Util.hs:87-88).
 * ---------------------------------------------------------------------------------------------------------- */

static float cat_at(const float *last, int nlast, const float *next, int idx) {
    return idx < nlast ? last[idx] : next[idx - nlast];
}

/* decimateCrossHighLevel FilterInternal.hs:398-402 (real); filterCrossHighLevel :405-408 is factor = 1 */
void o_decimateCrossR(int factor, int numCoeffs, const float *coeffs, int num, const float *last, int nlast,
                      const float *next, int nnext, float *out) {
    for (int o = 0; o < num; o++) {
        int   i = o * factor, avail = nlast - i + nnext;
        int   n = numCoeffs < avail ? numCoeffs : avail;
        float accum = 0;
        for (int k = 0; k < n; k++) accum = accum + cat_at(last, nlast, next, i + k) * coeffs[k];
        out[o] = accum;
    }
}

void o_decimateCrossC(int factor, int numCoeffs, const float *coeffs, int num, const float *last, int nlast,
                      const float *next, int nnext, float *out) {
    for (int o = 0; o < num; o++) {
        int   i = o * factor, avail = nlast - i + nnext;
        int   n = numCoeffs < avail ? numCoeffs : avail;
        float re = 0, im = 0;
        for (int k = 0; k < n; k++) {
            int idx = i + k;
            const float *s = idx < nlast ? last + 2 * idx : next + 2 * (idx - nlast);
            re = re + s[0] * coeffs[k];
            im = im + s[1] * coeffs[k];
        }
        out[2 * o]     = re;
        out[2 * o + 1] = im;
    }
}

/* resampleHighLevel FilterInternal.hs:253-265 / resampleCrossHighLevel :411-423.  Pass nlast = 0 and
 * last = NULL for the single-buffer form.  Taps used: coeffs[filterOffset + l*interpolation].  Returns the final
 * filterOffset. */
int o_resampleCrossR(int interpolation, int decimation, int numCoeffs, const float *coeffs, int filterOffset,
                     int count, const float *last, int nlast, const float *next, int nnext, float *out) {
    int inputOffset = 0;
    for (int i = 0; i < count; i++) {
        float accum = 0;
        int   avail = nlast + nnext - inputOffset;
        for (int l = 0, j = filterOffset; j < numCoeffs && l < avail; l++, j += interpolation)
            accum = accum + cat_at(last, nlast, next, inputOffset + l) * coeffs[j];
        out[i] = accum;
        int d = decimation - filterOffset - 1;
        /* divMod on non-negative operands here (decimation > filterOffset) */
        inputOffset += d / interpolation + 1;
        filterOffset = interpolation - 1 - d % interpolation;
    }
    return filterOffset;
}

int o_resampleCrossC(int interpolation, int decimation, int numCoeffs, const float *coeffs, int filterOffset,
                     int count, const float *last, int nlast, const float *next, int nnext, float *out) {
    int inputOffset = 0;
    for (int i = 0; i < count; i++) {
        float re = 0, im = 0;
        int   avail = nlast + nnext - inputOffset;
        for (int l = 0, j = filterOffset; j < numCoeffs && l < avail; l++, j += interpolation) {
            int idx = inputOffset + l;
            const float *s = idx < nlast ? last + 2 * idx : next + 2 * (idx - nlast);
            re = re + s[0] * coeffs[j];
            im = im + s[1] * coeffs[j];
        }
        out[2 * i]     = re;
        out[2 * i + 1] = im;
        int d = decimation - filterOffset - 1;
        inputOffset += d / interpolation + 1;
        filterOffset = interpolation - 1 - d % interpolation;
    }
    return filterOffset;
}

/* ------------------------------------------------------------------------------------------------------------
 * Converts and scale
 * ---------------------------------------------------------------------------------------------------------- */

/* convertC / convertCSSE / convertCAVX convert.c:15-50: (float(u8) - 128) * (1/128); all three agree bit-for-bit
 * (exact in binary32).  num counts BYTES, not IQ pairs (Util.hs:133). */
void o_convertC(int num, const uint8_t *in, float *out) {
    for (int i = 0; i < num; i++) out[i] = ((float)in[i] - 128.0f) * (1.0f / 128.0f);
}

/* convertCBladeRF / SSE / AVX convert.c:52-85: float(i16) * (1/2048) */
void o_convertCBladeRF(int num, const int16_t *in, float *out) {
    for (int i = 0; i < num; i++) out[i] = (float)in[i] * (1.0f / 2048.0f);
}

/* convertBladeRFTransmit convert.c:87-101.  (int16_t)val truncates toward zero; the clamps can never fire for
 * in-range values but are kept in the reference's order. */
void o_convertBladeRFTransmit(int num, const float *in, int16_t *out) {
    for (int i = 0; i < num; i++) {
        float val = in[i];
        val       = val + 1;
        val       = val * 2048;
        int16_t res = (int16_t)val;
        res         = res - 2048;
        if (res > 2047) res = 2047;
        if (res < -2048) res = -2048;
        out[i] = res;
    }
}

/* scale / scaleSSE / scaleAVX scale.c:15-36 */
void o_scale(int num, float factor, const float *in_buf, float *out_buf) {
    for (int i = 0; i < num; i++) out_buf[i] = in_buf[i] * factor;
}

/* ------------------------------------------------------------------------------------------------------------
 * FM discriminator (Demod.hs:21-46).  Haskell: phase (sample * conjugate last).
 *   (a:+b) * (c:+d) = (a*c - b*d) :+ (a*d + b*c) with conjugate last = lr :+ (-li):
 *      re = sr*lr - si*(-li)   im = sr*(-li) + si*lr       -- each product rounded, then one add/sub
 *   phase (0:+0) = 0 ; phase (x:+y) = atan2 y x            -- Data.Complex
 *   atan2 is the RealFloat class default (GHC.Float) built on atan = libm atanf for Float.
 * ---------------------------------------------------------------------------------------------------------- */

static int is_neg_zero(float v) { return v == 0.0f && signbit(v); }

static float hs_atan2f(float y, float x) {
    const float pi = 3.14159265358979323846f; /* pi :: Float */
    if (x > 0) return atanf(y / x);
    if (x == 0 && y > 0) return pi / 2;
    if (x < 0 && y > 0) return pi + atanf(y / x);
    if ((x <= 0 && y < 0) || (x < 0 && is_neg_zero(y)) || (is_neg_zero(x) && is_neg_zero(y)))
        return -hs_atan2f(-y, x);
    if (y == 0 && (x < 0 || is_neg_zero(x))) return pi;
    if (x == 0 && y == 0) return y;
    return x + y;
}

/* fmDemodVec Demod.hs:32-36.  last = (re, im) of the previous buffer's final sample (0,0 at stream start,
 * Demod.hs:41).  in: num interleaved complex samples; out: num floats. */
void o_fmDemod(int num, float lastRe, float lastIm, const float *in, float *out) {
    for (int i = 0; i < num; i++) {
        float sr = in[2 * i], si = in[2 * i + 1];
        float nli = -lastIm;
        float re = sr * lastRe - si * nli;
        float im = sr * nli + si * lastRe;
        out[i] = (re == 0 && im == 0) ? 0.0f : hs_atan2f(im, re);
        lastRe = sr;
        lastIm = si;
    }
}
