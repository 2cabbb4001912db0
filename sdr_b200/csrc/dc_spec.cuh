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
    int          ch, k1, k2;                 // chunk length, cheap warm-up, exact warm-up: multiples of 32
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
// the conversion instruction).  The float<->double conversions run on the XU pipe (about one warp instruction per 23
// cycles per scheduler) and looked like the bound of the first, per-lane version of the kernel (three per step).  Measured
// with the coalesced kernel the conversions win all the same: the ~8 integer instructions of a widening lengthen the
// dependent chain more than the conversion does (2^28 samples: 215 Gsamples/s widening both operands, 237 widening the
// difference only, 302 with conversions, 312 when the cheap warm-up converts too), so the default arithmetic flavour is
// DC_NATIVE_ALL and this function serves the
// other flavours, kept selectable for measurement (SDR_B200_DC_MODE).  == (double)f for all 2^32 bit patterns
// (tests/test_dc_speculation.py::test_widen_exhaustive_sample).
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
// one rounding to float on the assignment).  The float -> double widenings are exact whichever way they are done, so
// the flavours below give identical bits and differ only in which pipe they load:
//   DC_WIDEN_BOTH    integer widening of the difference and of y[n-1]: one XU conversion per step (the rounding)
//   DC_WIDEN_DIFF    integer widening of the difference only (it is off the dependent chain); y[n-1] by conversion
//   DC_NATIVE        both by conversion instructions: fewest instructions and the shortest dependent chain -- the
//                    serial kernel and the repair walk, where one lane runs alone, use it
//   DC_NATIVE_ALL    like DC_NATIVE, and the cheap warm-up converts too instead of widening
enum { DC_WIDEN_BOTH = 0, DC_WIDEN_DIFF = 1, DC_NATIVE = 2, DC_NATIVE_ALL = 3 };
template <int MODE> SDR_HD float dc_exact(float x, float last_sample, float last_output) {
#if defined(__CUDA_ARCH__)
    const float  d = __fsub_rn(x, last_sample);
    const double dd = (MODE >= DC_NATIVE) ? (double)d : dc_widen(d);
    const double yd = (MODE == DC_WIDEN_BOTH) ? dc_widen(last_output) : (double)last_output;
    return __double2float_rn(__dadd_rn(dd, __dmul_rn(0.997, yd)));
#else
    volatile float  d = x - last_sample;                 // volatile: no contraction, no excess precision
    volatile double p = 0.997 * ((MODE == DC_WIDEN_BOTH) ? dc_widen(last_output) : (double)last_output);
    volatile double s = ((MODE >= DC_NATIVE) ? (double)d : dc_widen(d)) + p;
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

// ---- one lane's walk over its chunk, in tiles of TILE samples ----------------------------------------------------------
// A lane visits positions [b0 - K1 - K2, b1) of the stream in TILE-sample tiles (chunk and warm-up lengths are multiples
// of the tile, so a tile lies in exactly one phase and every lane of a warp is in the same phase at the same time):
//   cheap warm-up  [b0 - K1 - K2, b0 - K2)   double state, one DFMA per sample
//   exact warm-up  [b0 - K2, b0)             the reference's arithmetic, nothing stored
//   owned          [b0, b1)                  the reference's arithmetic, outputs stored
// Tiles before the start of the stream are skipped; the lane then starts at position 0 from the true state.
// TILE (32 or 64 samples) is the granularity of the walk and of the staged copies: chunk and warm-up lengths are
// multiples of it.

struct DcLane {
    long long b0, b1;    // owned range
    long long pos;       // position of the next tile (negative: before the stream start)
    float     l, o;      // previous sample, current output (exact phases)
    double    a, ld;     // cheap phase: state and previous sample, widened
    bool      started;
};

SDR_HD void dc_lane_init(const DcArgs &A, long long c, DcLane &L) {
    L.b0 = c * A.ch;
    L.b1 = (L.b0 + A.ch < A.n) ? L.b0 + A.ch : A.n;
    if (c >= A.chunks) L.b1 = 0;      // no such chunk (tail of the last warp): every tile is skipped, warm-up included
    L.pos = L.b0 - A.k1 - A.k2;
    L.l = 0.0f; L.o = 0.0f; L.a = 0.0; L.ld = 0.0;
    L.started = false;
}
template <int TILE> SDR_HD long long dc_lane_tiles(const DcArgs &A) { return (A.k1 + A.k2 + A.ch) / TILE; }

// One tile.  `rd.get4(q)` delivers samples in[pos + 4q .. pos + 4q + 3], `wr.put4(q, v)` takes the four outputs of the
// same positions (owned tiles only; positions past the end of the stream carry the last value and are masked by the
// writer).  Returns true when the tile produced outputs.  The accessors keep only four samples live at a time: on the
// device they are shared-memory rows, on the host (and in the unaligned kernel) plain memory.
template <int MODE, int TILE, typename Rd, typename Wr>
SDR_HD bool dc_lane_tile(const DcArgs &A, long long c, DcLane &L, const Rd &rd, Wr &wr) {
    const long long pos = L.pos;
    L.pos += TILE;
    if (pos < 0 || pos >= L.b1) return false;
    if (!L.started) {
        L.started = true;
        if (pos == 0) {   // the true state before in[0]
            L.l = A.state_in ? A.state_in[0] : A.last_sample;
            L.o = A.state_in ? A.state_in[1] : A.last_output;
            L.a = dc_widen(L.o);
        } else {          // speculation: y = 0 at the start of the warm-up
            L.l = A.in[pos - 1];
            L.o = 0.0f; L.a = 0.0;
        }
        L.ld = dc_widen(L.l);
    }
    if (pos == L.b0 && c > 0) A.spec[c] = dc_bits(L.o);   // what this lane believes y[b0 - 1] to be
    if (pos < L.b0 - A.k2) {
        double a = L.a, ld = L.ld;
        float  last = L.l;
#pragma unroll
        for (int q = 0; q < TILE / 4; q++) {
            const float4 v = rd.get4(q);
            double xd;
            xd = (MODE == DC_NATIVE_ALL) ? (double)v.x : dc_widen(v.x); a = dc_cheap(xd, ld, a); ld = xd;
            xd = (MODE == DC_NATIVE_ALL) ? (double)v.y : dc_widen(v.y); a = dc_cheap(xd, ld, a); ld = xd;
            xd = (MODE == DC_NATIVE_ALL) ? (double)v.z : dc_widen(v.z); a = dc_cheap(xd, ld, a); ld = xd;
            xd = (MODE == DC_NATIVE_ALL) ? (double)v.w : dc_widen(v.w); a = dc_cheap(xd, ld, a); ld = xd;
            last = v.w;
        }
        L.a = a; L.ld = ld; L.l = last;
        L.o = (float)a;   // round-to-nearest on host and device; the exact phase continues from here
        return false;
    }
    float l = L.l, o = L.o;
    if (pos < L.b0) {
#pragma unroll
        for (int q = 0; q < TILE / 4; q++) {
            const float4 v = rd.get4(q);
            o = dc_exact<MODE>(v.x, l, o); o = dc_exact<MODE>(v.y, v.x, o); o = dc_exact<MODE>(v.z, v.y, o); o = dc_exact<MODE>(v.w, v.z, o);
            l = v.w;
        }
        L.l = l; L.o = o;
        return false;
    }
    const int m = (L.b1 - pos < TILE) ? (int)(L.b1 - pos) : TILE;   // < 32 only at the end of the stream
#pragma unroll
    for (int q = 0; q < TILE / 4; q++) {
        const float4 v = rd.get4(q);
        float4       y;
        if (4 * q + 0 < m) { o = dc_exact<MODE>(v.x, l, o); l = v.x; } y.x = o;
        if (4 * q + 1 < m) { o = dc_exact<MODE>(v.y, l, o); l = v.y; } y.y = o;
        if (4 * q + 2 < m) { o = dc_exact<MODE>(v.z, l, o); l = v.z; } y.z = o;
        if (4 * q + 3 < m) { o = dc_exact<MODE>(v.w, l, o); l = v.w; } y.w = o;
        wr.put4(q, y);
    }
    L.l = l; L.o = o;
    if (pos + TILE >= L.b1) A.fin[c] = dc_bits(o);
    return true;
}

// accessors over plain memory, bounds-checked against the end of the stream
struct DcMemReader {
    const float *in; long long pos, n;
    SDR_HD float  at(long long i) const { return i < n ? in[i] : 0.0f; }
    SDR_HD float4 get4(int q) const {
        const long long i = pos + 4 * q;
        return make_float4(at(i), at(i + 1), at(i + 2), at(i + 3));
    }
};
struct DcMemWriter {
    float *out; long long pos, end;
    SDR_HD void put4(int q, const float4 &v) {
        const long long i = pos + 4 * q;
        if (i < end) out[i] = v.x;
        if (i + 1 < end) out[i + 1] = v.y;
        if (i + 2 < end) out[i + 2] = v.z;
        if (i + 3 < end) out[i + 3] = v.w;
    }
};

// The whole walk of chunk c straight out of / into global memory: the kernel for buffers that are not 16-byte aligned,
// and what the CPU tests run (tests/emul/dc_emul.cpp).  The aligned kernel (k_dc_spec_tiles) makes the same calls with
// rows staged through shared memory by coalesced asynchronous copies.
template <int MODE, int TILE = 32> SDR_HD void dc_chunk(const DcArgs &A, long long c) {
    DcLane L;
    dc_lane_init(A, c, L);
    const long long tiles = dc_lane_tiles<TILE>(A);
    for (long long t = 0; t < tiles; t++) {
        const DcMemReader rd = {A.in, L.pos, A.n};
        DcMemWriter       wr = {A.out, L.pos, L.b1};
        dc_lane_tile<MODE, TILE>(A, c, L, rd, wr);
    }
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
        o = dc_exact<DC_NATIVE>(x, l, o); l = x;
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
