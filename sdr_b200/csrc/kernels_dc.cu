// kernels_dc.cu -- dcBlocker (c_sources/filter.c:152-161) / dcBlockingFilter (hs_sources/SDR/Filter.hs:730-739) on the
// device: the serial one-lane kernel for short vectors and the speculative chunk-parallel evaluation of dc_spec.cuh for
// long ones.  Both are bit-exact; see dc_spec.cuh for why the parallel one is.
#include "common.cuh"
#include "dc_spec.cuh"

#include <cstdlib>

namespace sdr {

#define SDR_LAUNCH_CHECK(c)                       \
    do {                                          \
        (c)->launches++;                          \
        SDR_CUDA(cudaGetLastError());             \
    } while (0)

// Serial form.  Each y[n] depends on the ROUNDED y[n-1]: one lane evaluates the recurrence, the warp only streams the
// data through shared memory in coalesced 1024-sample pieces.
__global__ void __launch_bounds__(32) k_dc_blocker(float last_sample, float last_output, const float *__restrict__ state_in,
                                                   const float *__restrict__ in, float *__restrict__ out, long long n,
                                                   float *__restrict__ final2) {
    __shared__ float buf[1024];
    if (state_in) { last_sample = state_in[0]; last_output = state_in[1]; }   // streaming form: state carried on the device
    for (long long base = 0; base < n; base += 1024) {
        int m = (int)((n - base) < 1024 ? (n - base) : 1024);
        for (int i = threadIdx.x; i < m; i += 32) buf[i] = in[base + i];
        __syncwarp();
        if (threadIdx.x == 0) {
            for (int i = 0; i < m; i++) {
                float x = buf[i];
                last_output = dc_exact<DC_NATIVE>(x, last_sample, last_output);
                last_sample = x;
                buf[i] = last_output;
            }
        }
        __syncwarp();
        for (int i = threadIdx.x; i < m; i += 32) out[base + i] = buf[i];
        __syncwarp();
    }
    if (threadIdx.x == 0) { final2[0] = last_sample; final2[1] = last_output; }
}

// Speculative pass, buffers with any 4-byte alignment: one lane per chunk, straight out of global memory.
__global__ void __launch_bounds__(128) k_dc_spec(DcArgs A) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c < A.chunks) dc_chunk<DC_NATIVE_ALL, 32>(A, c);
}

// Speculative pass, 16-byte aligned buffers.  A warp owns 32 consecutive chunks, one per lane, and all its lanes sit at
// the same offset inside their chunks.  Per tile the warp copies one TILE-sample piece (256 bytes by default) of each of
// its 32 chunks into shared memory with coalesced 16-byte asynchronous copies (TILE/4 lanes per piece, ahead of the
// arithmetic: a tile is ~4500 cycles of dependent arithmetic per lane, which is what hides DRAM), every lane then reads
// ITS row with conflict-free LDS.128, and the outputs go back the same way.  The per-lane version above reads 32
// separate sectors per warp instruction, each in its own DRAM page: ncu showed it waiting on memory (long_scoreboard
// 5.4 warps per issue at 19 % of DRAM bandwidth).
namespace {
__device__ __forceinline__ void dc_cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void dc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void dc_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
struct DcSmemReader {
    const float4 *row;
    __device__ __forceinline__ float4 get4(int q) const { return row[q]; }
};
struct DcSmemWriter {
    float4 *row;
    __device__ __forceinline__ void put4(int q, const float4 &v) { row[q] = v; }
};
}  // namespace

// TILE samples (128 or 256 bytes) of each of the warp's 32 chunks per staged tile; WARPS warps per block.  A row holds
// TILE + 4 floats: an odd number of 16-byte chunks, so the eight lanes of an LDS.128 phase hit eight distinct bank groups.
// NBUF input tiles per warp, NBUF - 1 of them in flight ahead of the arithmetic.
template <int MODE, int TILE, int WARPS, int NBUF>
__global__ void __launch_bounds__(32 * WARPS) k_dc_spec_tiles(DcArgs A) {
    constexpr int ROW = TILE + 4, TILE_FLOATS = 32 * ROW;
    constexpr int PIECES = TILE / 4;          // 16-byte pieces per row
    constexpr int ROWS_PER_INSTR = 32 / PIECES, INSTRS = 32 / ROWS_PER_INSTR;
    static_assert((ROW / 4) % 2 == 1 && 32 % PIECES == 0, "row stride / copy geometry");
    __shared__ __align__(16) float s_in[WARPS][NBUF][TILE_FLOATS];
    __shared__ __align__(16) float s_out[WARPS][TILE_FLOATS];
    const int       lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long c_first = ((long long)blockIdx.x * WARPS + warp) * 32;   // this warp's first chunk
    if (c_first >= A.chunks) return;
    const long long c = c_first + lane;
    const long long tiles = dc_lane_tiles<TILE>(A);
    const long long rel0 = -(long long)A.k1 - A.k2;                  // offset of tile 0 from the start of a chunk
    const int       sub_row = lane / PIECES, piece = lane % PIECES;  // this lane's part in the cooperative copies
    DcLane L;
    dc_lane_init(A, c, L);

    // copy tile t of all 32 chunks: instruction j moves one 16-byte piece of ROWS_PER_INSTR rows
    auto issue = [&](long long t) {
        if (t < tiles) {
            float *buf = s_in[warp][t % NBUF];
#pragma unroll
            for (int j = 0; j < INSTRS; j++) {
                const int       row = ROWS_PER_INSTR * j + sub_row;
                const long long cr = c_first + row;
                const long long e = cr * A.ch + rel0 + t * TILE + piece * 4;   // first element of the piece
                long long       left = A.n - e;                                // elements of it inside the stream
                if (cr < A.chunks && e >= 0 && left > 0)
                    dc_cp_async16((uint32_t)__cvta_generic_to_shared(buf + row * ROW + piece * 4), A.in + e,
                                  left >= 4 ? 16u : (uint32_t)left * 4u);
            }
        }
        dc_cp_commit();
    };
#pragma unroll
    for (int t = 0; t < NBUF - 1; t++) issue(t);
    for (long long t = 0; t < tiles; t++) {
        issue(t + NBUF - 1);
        dc_cp_wait<NBUF - 1>();
        __syncwarp();
        const DcSmemReader rd = {reinterpret_cast<const float4 *>(s_in[warp][t % NBUF] + lane * ROW)};
        DcSmemWriter       wr = {reinterpret_cast<float4 *>(s_out[warp] + lane * ROW)};
        dc_lane_tile<MODE, TILE>(A, c, L, rd, wr);
        // owned tiles: every lane of the warp is in its owned range at the same time (the phase depends on the offset only)
        if (rel0 + t * TILE >= 0) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < INSTRS; j++) {
                const int       rw = ROWS_PER_INSTR * j + sub_row;
                const long long cr = c_first + rw;
                const long long e = cr * A.ch + rel0 + t * TILE + piece * 4;
                const long long left = A.n - e;
                if (cr < A.chunks && left > 0) {
                    const float4 v = *reinterpret_cast<const float4 *>(s_out[warp] + rw * ROW + piece * 4);
                    if (left >= 4) *reinterpret_cast<float4 *>(A.out + e) = v;
                    else { A.out[e] = v.x; if (left > 1) A.out[e + 1] = v.y; if (left > 2) A.out[e + 2] = v.z; }
                }
            }
        }
        __syncwarp();   // the input buffer of tile t is refilled by the next iteration's copies
    }
}

// Check + repair.  All threads compare spec[c] with fin[c-1] and build the bitmap of missed chunks; if there is none
// (the usual case) the kernel only publishes the final state, otherwise thread 0 repairs them in stream order.
__global__ void __launch_bounds__(1024) k_dc_repair(DcArgs A) {
    int any = 0;
    for (long long c0 = (long long)(threadIdx.x / 32) * 32; c0 < A.chunks; c0 += blockDim.x) {
        const long long c = c0 + (threadIdx.x & 31);
        const bool miss = c > 0 && c < A.chunks && dc_missed(A, c, A.fin[c - 1]);
        const unsigned m = __ballot_sync(0xffffffffu, miss);
        if ((threadIdx.x & 31) == 0) A.fail_bits[c0 / 32] = m;
        any |= (m != 0);
    }
    any = __syncthreads_or(any);   // also orders the bitmap writes before thread 0's reads
    if (threadIdx.x != 0) return;
    if (any) { dc_repair(A); return; }
    A.stats[0] += 1; A.stats[1] += (unsigned long long)A.chunks;
    if (A.final2) {
        const float fs = A.in[A.n - 1], fo = dc_float(A.fin[A.chunks - 1]);
        A.final2[0] = fs; A.final2[1] = fo;
    }
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    if (!e || !*e) return dflt;
    int v = atoi(e);
    return v >= 0 ? v : dflt;
}

// Tuning (dc_spec.cuh): chunk length 0 = automatic, warm-up lengths.  Every lane pays K1 + K2 warm-up samples per chunk, so
// long chunks waste less arithmetic and re-read less input, while short ones give more lanes to hide the ~70-cycle
// dependent step.  Measured (tools/dc_sweep.py, profiles/r01b_dc_sweep.txt, B200), 2^28 samples: 2048 -> 127, 4096 -> 208,
// 6144 -> 272, 8192 -> 303, 12288 -> 239, 16384 -> 224 Gsamples/s; 2^27: 4096 -> 208, 8192 -> 178: the optimum sits at about
// 32768 lanes = 7 warps per SM.  Automatic = that many lanes, between 2048 and 8192 samples.
static void dc_tuning(const Ctx *c, long long n, int *ch, int *k1, int *k2) {
    int want = c->dc_chunk > 0 ? c->dc_chunk : env_int("SDR_B200_DC_CHUNK", 0);
    *k1 = c->dc_k1 >= 0 ? c->dc_k1 : env_int("SDR_B200_DC_K1", 6144);
    *k2 = c->dc_k2 >= 0 ? c->dc_k2 : env_int("SDR_B200_DC_K2", 4096);
    if (want <= 0) {
        long long lanes = (long long)c->sm_count * 224;
        long long per = (n + lanes - 1) / lanes;
        if (per < 2048) per = 2048;
        if (per > 8192) per = 8192;
        want = (int)per;
    }
    *ch = (want + 63) & ~63;   // multiples of the larger tile
    *k1 = (*k1 + 63) & ~63;
    *k2 = (*k2 + 63) & ~63;
}

static bool overlap(const void *a, const void *b, size_t bytes) {
    const char *p = (const char *)a, *q = (const char *)b;
    return p < q + bytes && q < p + bytes;
}

static int dc_run(Ctx *c, float last_sample, float last_output, const float *d_state, const float *d_in, float *d_out, long long n,
                  float *d_final2) {
    if (n <= 0) return SDR_OK;
    SDR_TRY(c->bind());
    const long long min_parallel = c->dc_min_parallel >= 0 ? c->dc_min_parallel : 65536;
    const bool aligned4 = (((uintptr_t)d_in | (uintptr_t)d_out) & 3) == 0;
    if (n < min_parallel || !aligned4 || overlap(d_in, d_out, (size_t)n * 4)) {   // in-place calls keep the serial kernel
        k_dc_blocker<<<1, 32, 0, c->s()>>>(last_sample, last_output, d_state, d_in, d_out, n, d_final2);
        SDR_LAUNCH_CHECK(c);
        c->dc_last_parallel = 0;
        return SDR_OK;
    }
    DcArgs A;
    A.in = d_in; A.out = d_out; A.n = n;
    A.last_sample = last_sample; A.last_output = last_output; A.state_in = d_state;
    dc_tuning(c, n, &A.ch, &A.k1, &A.k2);
    A.chunks = (n + A.ch - 1) / A.ch;
    const size_t words = (size_t)((A.chunks + 31) / 32);
    const size_t need = 64 + (size_t)A.chunks * 8 + words * 4;
    SDR_TRY(c->ensure_dc_scratch(need));
    char *base = (char *)c->d_dc_scratch;
    if (c->dc_scratch_regrown) {
        SDR_CUDA(cudaMemsetAsync(base, 0, 64, c->s()));   // the counters start at zero (and restart when the block moves)
        c->dc_scratch_regrown = false;
    }
    A.stats = (unsigned long long *)base;
    A.spec = (uint32_t *)(base + 64);
    A.fin = A.spec + A.chunks;
    A.fail_bits = A.fin + A.chunks;
    A.final2 = d_final2;
    const bool vec = (((uintptr_t)d_in | (uintptr_t)d_out) & 15) == 0;
    static const int mode = env_int("SDR_B200_DC_MODE", DC_NATIVE_ALL);   // measurement knobs: every flavour gives identical bits
    static const int tile = env_int("SDR_B200_DC_TILE", 64);   // 256-byte tiles: 338 vs 311 Gsamples/s at 2^28 samples
    const int g2 = (int)((A.chunks + 63) / 64), g1 = (int)((A.chunks + 31) / 32);
    if (vec && tile == 64)                  k_dc_spec_tiles<DC_NATIVE_ALL, 64, 1, 2><<<g1, 32, 0, c->s()>>>(A);
    else if (vec && mode == DC_WIDEN_BOTH)  k_dc_spec_tiles<DC_WIDEN_BOTH, 32, 2, 3><<<g2, 64, 0, c->s()>>>(A);
    else if (vec && mode == DC_WIDEN_DIFF)  k_dc_spec_tiles<DC_WIDEN_DIFF, 32, 2, 3><<<g2, 64, 0, c->s()>>>(A);
    else if (vec && mode == DC_NATIVE)      k_dc_spec_tiles<DC_NATIVE, 32, 2, 3><<<g2, 64, 0, c->s()>>>(A);
    else if (vec)                           k_dc_spec_tiles<DC_NATIVE_ALL, 32, 2, 3><<<g2, 64, 0, c->s()>>>(A);
    else     k_dc_spec<<<(int)((A.chunks + 127) / 128), 128, 0, c->s()>>>(A);
    SDR_LAUNCH_CHECK(c);
    k_dc_repair<<<1, 1024, 0, c->s()>>>(A);
    SDR_LAUNCH_CHECK(c);
    c->dc_last_parallel = 1;
    return SDR_OK;
}

int launch_dc_blocker(Ctx *c, float last_sample, float last_output, const float *d_in, float *d_out, long long n,
                      float *d_final2) {
    return dc_run(c, last_sample, last_output, nullptr, d_in, d_out, n, d_final2);
}
// d_state: (lastSample, lastOutput) on the device, read before and updated after the block (dcBlockingFilter's pMapAccum)
int launch_dc_blocker_carry(Ctx *c, float *d_state, const float *d_in, float *d_out, long long n) {
    return dc_run(c, 0.0f, 0.0f, d_state, d_in, d_out, n, d_state);
}

}  // namespace sdr

using namespace sdr;

extern "C" {

int sdr_dev_dc_blocker(sdr_ctx_t *ctx, float last_sample, float last_output, const float *d_in, float *d_out, long long n,
                       float *d_final2) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || n < 0 || (n && (!d_in || !d_out)) || !d_final2) return set_error(SDR_EINVAL, "sdr_dev_dc_blocker: bad argument");
    if (n == 0) {
        const float st[2] = {last_sample, last_output};
        SDR_TRY(c->bind());
        SDR_CUDA(cudaMemcpyAsync(d_final2, st, 8, cudaMemcpyHostToDevice, c->s()));
        SDR_CUDA(cudaStreamSynchronize(c->s()));
        return SDR_OK;
    }
    return launch_dc_blocker(c, last_sample, last_output, d_in, d_out, n, d_final2);
}

int sdr_dc_blocker_tuning(sdr_ctx_t *ctx, int chunk, int cheap_warmup, int exact_warmup, long long min_parallel) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return set_error(SDR_EINVAL, "sdr_dc_blocker_tuning: null context");
    if (chunk > 0 && chunk < 64) return set_error(SDR_EINVAL, "sdr_dc_blocker_tuning: chunk %d < 64", chunk);
    c->dc_chunk = chunk; c->dc_k1 = cheap_warmup; c->dc_k2 = exact_warmup; c->dc_min_parallel = min_parallel;
    return SDR_OK;
}

int sdr_dc_blocker_stats(sdr_ctx_t *ctx, long long stats[4], int *last_parallel) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !stats) return set_error(SDR_EINVAL, "sdr_dc_blocker_stats: bad argument");
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (last_parallel) *last_parallel = c->dc_last_parallel;
    if (!c->d_dc_scratch) return SDR_OK;
    SDR_TRY(c->bind());
    SDR_CUDA(cudaStreamSynchronize(c->s()));
    unsigned long long h[4];
    SDR_CUDA(cudaMemcpy(h, c->d_dc_scratch, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; i++) stats[i] = (long long)h[i];
    return SDR_OK;
}

}  // extern "C"
