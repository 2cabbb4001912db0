// dc_spec.cuh -- speculative, chunk-parallel and still BIT-EXACT evaluation of the reference's dcBlocker
// (c_sources/filter.c:152-161, driven by dcBlockingFilter, hs_sources/SDR/Filter.hs:730-739).
//
//   y[n] = fl32( fl64( (double)fl32(x[n] - x[n-1]) + fl64(0.997 * (double)y[n-1]) ) )
//
// Every y[n] depends on the ROUNDED y[n-1], so a prefix scan in real arithmetic is only norm-accurate.  What makes a
// parallel evaluation exact is that the recurrence is a contraction (0.997 < 1) followed by a rounding onto the float
// grid: two trajectories driven by the same input approach each other geometrically and, once within an ulp, snap onto
// the same float -- from that sample on they are bit-identical for ever.  So the stream is cut into chunks, one lane
// per chunk:
//   1. the lane starts K1 + K2 samples before its chunk from y = 0 (or from the true state, near the stream start);
//      K1 samples in cheap arithmetic (one double FMA per sample: it only has to land within a few float ulps of the
//      true trajectory), then K2 samples in the exact arithmetic so the trajectories merge;
//   2. it remembers the value it reached just before its chunk (`spec`), evaluates its chunk exactly and stores the
//      outputs and its final value (`fin`);
//   3. chunk c is exact iff spec[c] == fin[c-1] bit for bit (chunk 0 starts from the true state).  Chunks that fail
//      are re-evaluated serially from fin[c-1], in stream order, until the new trajectory meets the stored one
//      (dc_repair); the result is therefore bit-exact for ANY input, speculation only decides the speed.
// Measured on the CPU model (tests/test_dc_speculation.py): after K1 = 6144 cheap steps the exact phase merges after
// 390 samples on average, 3000 worst of 2100 trials on white noise; K2 = 4096.  Input that is exactly constant for
// > 30k samples parks the true trajectory on a denormal fixed point (0.997*y rounds back to y for |y| < 167 denormal
// steps) that speculation from 0 never reaches: every chunk is then repaired and the speed is the serial kernel's.
//
// The functions are __host__ __device__ so that the CPU tests drive the very code the kernels run (tests/emul/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SDR_HD __host__ __device__ __forceinline__
#else
#define SDR_HD inline
#endif

namespace sdr {

struct DcArgs {
    const float *in;
    float       *out;
    long long    n;
    float        last_sample, last_output;   // state before in[0] ...
    const float *state_in;                   // ... or, when non-null, (lastSample, lastOutput) read from here
    int          ch, k1, k2;                 // chunk length, cheap warm-up, exact warm-up: multiples of 8
    long long    chunks;                     // ceil(n / ch)
    uint32_t    *spec, *fin;                 // [chunks] value reached before the chunk / at its end (float bits)
    uint32_t    *fail_bits;                  // [(chunks + 31) / 32] bitmap of chunks whose speculation missed
    unsigned long long *stats;               // [0] launches [1] chunks [2] repaired chunks [3] repaired samples
    float       *final2;                     // (finalSample, finalOutput)
};

SDR_HD uint32_t dc_bits(float v) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(v);
#else
    uint32_t u; memcpy(&u, &v, 4); return u;
#endif
}
SDR_HD float dc_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float v; memcpy(&v, &u, 4); return v;
#endif
}

// Exact float -> double widening with integer operations (normal numbers and zeros; denormals, infinities and NaN take
// the conversion instruction).  On the device the float<->double conversions run on the XU pipe at about one warp
// instruction per 23 cycles and were the bound of the first version of the speculative kernel (ncu: XU 101 % busy, three
// conversions per step); the recurrence itself needs only the one rounding double -> float.  Checked against (double)f
// for all 2^32 bit patterns (tests/test_dc_speculation.py::test_widen_exhaustive_sample).
SDR_HD double dc_widen(float f) {
    const uint32_t b = dc_bits(f);
    const uint32_t mag = b & 0x7fffffffu;
    if (mag != 0 && mag - 0x00800000u >= 0x7f000000u) return (double)f;   // denormal, inf, NaN
    uint32_t hi = (mag >> 3) + 0x38000000u;   // exponent rebias 127 -> 1023, top 20 mantissa bits
    hi = (mag == 0 ? 0u : hi) | (b & 0x80000000u);
    const uint32_t lo = b << 29;              // low 3 mantissa bits
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    const uint64_t u = ((uint64_t)hi << 32) | lo;
    double d; memcpy(&d, &u, 8); return d;
#endif
}

// one exact step: returns y[n] from x[n], x[n-1], y[n-1] (filter.c:155: float difference, double product and sum,
// one rounding to float on the assignment)
SDR_HD float dc_exact(float x, float last_sample, float last_output) {
#if defined(__CUDA_ARCH__)
    return __double2float_rn(__dadd_rn(dc_widen(__fsub_rn(x, last_sample)), __dmul_rn(0.997, dc_widen(last_output))));
#else
    volatile float  d = x - last_sample;                 // volatile: no contraction, no excess precision
    volatile double p = 0.997 * dc_widen(last_output);
    volatile double s = dc_widen(d) + p;
    return (float)s;
#endif
}
// one cheap step on a double state (warm-up only; never stored): xd, ld = the sample and its predecessor widened
SDR_HD double dc_cheap(double xd, double ld, double a) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(0.997, a, __dsub_rn(xd, ld));
#else
    return __builtin_fma(0.997, a, xd - ld);
#endif
}

struct DcGroup { float v[8]; };

template <bool VEC> SDR_HD DcGroup dc_load8(const float *in, long long pos) {
    DcGroup g;
    if (VEC) {   // in is 16-byte aligned and pos a multiple of 8
        const float4 a = *reinterpret_cast<const float4 *>(in + pos), b = *reinterpret_cast<const float4 *>(in + pos + 4);
        g.v[0] = a.x; g.v[1] = a.y; g.v[2] = a.z; g.v[3] = a.w; g.v[4] = b.x; g.v[5] = b.y; g.v[6] = b.z; g.v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) g.v[i] = in[pos + i];
    }
    return g;
}
template <bool VEC> SDR_HD void dc_store8(float *out, long long pos, const DcGroup &g) {
    if (VEC) {
        *reinterpret_cast<float4 *>(out + pos)     = make_float4(g.v[0], g.v[1], g.v[2], g.v[3]);
        *reinterpret_cast<float4 *>(out + pos + 4) = make_float4(g.v[4], g.v[5], g.v[6], g.v[7]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) out[pos + i] = g.v[i];
    }
}

// Speculative evaluation of chunk c (steps 1 and 2 above).  Every group boundary (w, e0, b0) is a multiple of 8, so the
// whole walk is one run of 8-sample groups with the next group always loaded before the current one is evaluated.
template <bool VEC> SDR_HD void dc_chunk(const DcArgs &A, long long c) {
    const long long b0 = c * A.ch;
    const long long b1 = (b0 + A.ch < A.n) ? b0 + A.ch : A.n;
    const long long full_end = b0 + ((b1 - b0) & ~7LL);   // end of the whole 8-sample groups (b1 except in the last chunk)
    const float s0 = A.state_in ? A.state_in[0] : A.last_sample;
    const float o0 = A.state_in ? A.state_in[1] : A.last_output;

    long long e0 = b0 - A.k2; if (e0 < 0) e0 = 0;          // exact warm-up  [e0, b0)
    long long w  = e0 - A.k1; if (w < 0) w = 0;            // cheap warm-up  [w, e0)
    if (c == 0) { e0 = 0; w = 0; }
    float  l = (w == 0) ? s0 : A.in[w - 1];
    double a = (w == 0) ? (double)o0 : 0.0;

    // the next two groups are always in flight before the current one is evaluated (a group is ~500 cycles of
    // dependent arithmetic, a DRAM access under load rather more)
    long long pos = w;
    DcGroup   cur = DcGroup(), nx1 = DcGroup(), nx2;
    if (pos < full_end) cur = dc_load8<VEC>(A.in, pos);
    if (pos + 8 < full_end) nx1 = dc_load8<VEC>(A.in, pos + 8);
    double ld = dc_widen(l);
    for (; pos < e0; pos += 8) {
        nx2 = nx1;
        if (pos + 16 < full_end) nx2 = dc_load8<VEC>(A.in, pos + 16);
#pragma unroll
        for (int i = 0; i < 8; i++) { const double xd = dc_widen(cur.v[i]); a = dc_cheap(xd, ld, a); ld = xd; }
        l = cur.v[7];
        cur = nx1; nx1 = nx2;
    }
    float o = (float)a;   // exact when nothing was warmed up cheaply (a == o0); (float) is round-to-nearest on both sides
    for (; pos < b0; pos += 8) {
        nx2 = nx1;
        if (pos + 16 < full_end) nx2 = dc_load8<VEC>(A.in, pos + 16);
#pragma unroll
        for (int i = 0; i < 8; i++) { o = dc_exact(cur.v[i], l, o); l = cur.v[i]; }
        cur = nx1; nx1 = nx2;
    }
    if (c > 0) A.spec[c] = dc_bits(o);
    for (; pos < full_end; pos += 8) {
        nx2 = nx1;
        if (pos + 16 < full_end) nx2 = dc_load8<VEC>(A.in, pos + 16);
        DcGroup y;
#pragma unroll
        for (int i = 0; i < 8; i++) { o = dc_exact(cur.v[i], l, o); l = cur.v[i]; y.v[i] = o; }
        dc_store8<VEC>(A.out, pos, y);
        cur = nx1; nx1 = nx2;
    }
#pragma unroll 1
    for (; pos < b1; pos++) {   // ragged end of the stream (last chunk only)
        const float x = A.in[pos];
        o = dc_exact(x, l, o); l = x;
        A.out[pos] = o;
    }
    A.fin[c] = dc_bits(o);
}

// chunk c missed iff the value it reached just before its first sample differs from the true one
SDR_HD bool dc_missed(const DcArgs &A, long long c, uint32_t true_prev) { return A.spec[c] != true_prev; }

// Step 3 for one chunk whose speculation missed: re-evaluate from the true state until the new trajectory meets the
// stored one.  Returns the true final value of the chunk (float bits) and adds to *samples what it rewrote.
SDR_HD uint32_t dc_repair_chunk(const DcArgs &A, long long c, uint32_t true_prev, unsigned long long *samples) {
    const long long b0 = c * A.ch;
    const long long b1 = (b0 + A.ch < A.n) ? b0 + A.ch : A.n;
    float l = A.in[b0 - 1], o = dc_float(true_prev);
    for (long long i = b0; i < b1; i++) {
        const float x = A.in[i];
        o = dc_exact(x, l, o); l = x;
        if (dc_bits(o) == dc_bits(A.out[i])) { *samples += (unsigned long long)(i - b0); return A.fin[c]; }   // merged: the rest stands
        A.out[i] = o;
    }
    *samples += (unsigned long long)(b1 - b0);
    A.fin[c] = dc_bits(o);
    return dc_bits(o);
}

// Serial tail of step 3, in stream order over the chunks flagged in fail_bits; a repair that changes a chunk's final
// value re-opens the check of its successor.  Also publishes the final state.  One thread.
SDR_HD void dc_repair(const DcArgs &A) {
    unsigned long long repaired = 0, samples = 0;
    const long long words = (A.chunks + 31) / 32;
    bool      carry = false;        // the previous chunk's final value changed: check this chunk against `prev`
    uint32_t  prev = 0;
    for (long long wi = 0; wi < words; wi++) {
        uint32_t bits = A.fail_bits[wi];
        if (!bits && !carry) continue;
        for (int b = 0; b < 32; b++) {
            const long long c = wi * 32 + b;
            if (c >= A.chunks) break;
            bool miss = (bits >> b) & 1u;
            if (carry) miss = (c > 0) && dc_missed(A, c, prev);
            carry = false;
            if (!miss || c == 0) continue;
            const uint32_t true_prev = A.fin[c - 1];
            const uint32_t before = A.fin[c];
            prev = dc_repair_chunk(A, c, true_prev, &samples);
            repaired++;
            carry = (prev != before);
        }
    }
    A.stats[0] += 1; A.stats[1] += (unsigned long long)A.chunks; A.stats[2] += repaired; A.stats[3] += samples;
    if (A.final2) {
        const float fs = A.in[A.n - 1], fo = dc_float(A.fin[A.chunks - 1]);   // read before the writes: final2 may alias state_in
        A.final2[0] = fs; A.final2[1] = fo;
    }
}

}  // namespace sdr
