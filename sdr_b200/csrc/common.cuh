// common.cuh -- shared declarations of the sdr_b200 native library (host + device).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/sdr_b200.h"

namespace sdr {

// ---- error plumbing -------------------------------------------------------------------------------------------
int set_error(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SDR_CUDA(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return ::sdr::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define SDR_TRY(expr)            \
    do {                         \
        int _s = (expr);         \
        if (_s != SDR_OK) return _s; \
    } while (0)

// ---- context --------------------------------------------------------------------------------------------------
struct Ctx {
    int          device      = 0;
    cudaStream_t stream      = nullptr;
    cudaStream_t side        = nullptr;  // halo exchange / ragged-tail overlap
    cudaStream_t override_st = nullptr;  // when set, kernel launchers enqueue here instead of `stream`
    cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t s() const { return override_st ? override_st : stream; }
    int          reserve_sms = 0;        // SMs the persistent kernels leave free (for a concurrent NCCL kernel)
    int          sm_count    = 0;
    int          arith       = SDR_ARITH_FAST;
    bool         fir_ffa     = false;    // real stride-1 filters: 2-parallel fast-FIR arithmetic (kernels_real.cu)
    long long    launches    = 0;
    // device staging for HOST-pointer calls (grown on demand; copies go straight from / to the caller's memory)
    void  *d_stage_in = nullptr; size_t d_stage_in_bytes = 0;
    void  *d_stage_out = nullptr; size_t d_stage_out_bytes = 0;
    void  *d_flush = nullptr; size_t d_flush_bytes = 0;
    // dcBlocker speculation (kernels_dc.cu): per-chunk scratch + counters, tuning overrides (0 / -1 = default)
    void  *d_dc_scratch = nullptr; size_t d_dc_scratch_bytes = 0; bool dc_scratch_regrown = false;
    int    dc_chunk = 0, dc_k1 = -1, dc_k2 = -1, dc_last_parallel = 0;
    long long dc_min_parallel = -1;
    int ensure_stage(size_t in_bytes, size_t out_bytes);
    int ensure_dc_scratch(size_t bytes);
    int bind() const;  // cudaSetDevice
};

// ---- a logically contiguous input made of up to two device segments (lastBuf ++ nextBuf) --------------------------
struct Seg2 {
    const void *a; long long na;  // first segment, na ELEMENTS (floats for real, float2 for complex)
    const void *b; long long nb;  // second segment
};

// ---- kernel launchers (kernels_generic.cu) ----------------------------------------------------------------------
// y[m] = sum_k c[k] x[m*D + k], m < num.  d_taps: T floats on device.  x = seg.a ++ seg.b, n_in = na + nb elements;
// elements beyond n_in read as zero.
int launch_fir_generic(Ctx *c, bool cplx, int T, int D, const float *d_taps, Seg2 seg, void *d_out, long long num);
// same math, unfused mul+add in a reference variant's lane order.  W = SIMD width in floats (1 scalar, 4 SSE, 8 AVX);
// layout 0 real / 1 complex duplicated-coefficient form / 2 complex "2" form; sym: T = half taps, pre-added samples.
int launch_fir_exact_fir(Ctx *c, bool cplx, int T, int D, int W, int layout, int sym, const float *d_taps, Seg2 seg,
                         void *d_out, long long num);
// polyphase group-table resampler (resample.c:34-142): output i uses group (g0+i)%ng at input offset
// ((g0+i)/ng)*sum_inc + prefix[(g0+i)%ng] - prefix[g0]; d_table [ng][row_stride]; d_prefix [ng]
int launch_resample_groups(Ctx *c, bool cplx, int taps_per_group, int row_stride, int g0, int ng, const int *d_prefix,
                           int sum_inc, const float *d_table, Seg2 seg, void *d_out, long long num);
int launch_resample_exact(Ctx *c, bool cplx, int taps_per_group, int row_stride, int W, int layout, int g0, int ng,
                          const int *d_prefix, int sum_inc, const float *d_table, Seg2 seg, void *d_out, long long num);
int launch_convert_u8(Ctx *c, const uint8_t *d_in, float *d_out, long long n);
int launch_convert_i16(Ctx *c, const int16_t *d_in, float *d_out, long long n);
int launch_convert_tx(Ctx *c, const float *d_in, int16_t *d_out, long long n);
int launch_scale(Ctx *c, float k, const float *d_in, float *d_out, long long n);
int launch_fm_demod(Ctx *c, float last_re, float last_im, const float *d_in, float *d_out, long long n);
int launch_fm_demod_carry(Ctx *c, const float *d_last, const float *d_in, float *d_out, long long n);
int launch_dc_blocker(Ctx *c, float last_sample, float last_output, const float *d_in, float *d_out, long long n,
                      float *d_final2);
int launch_dc_blocker_carry(Ctx *c, float *d_state, const float *d_in, float *d_out, long long n);
int launch_synth_noise(Ctx *c, float *d_out, long long n, long long first, uint32_t seed);
int launch_synth_bytes(Ctx *c, uint8_t *d_out, long long n, long long first, uint32_t seed);
int launch_checksum32(Ctx *c, const uint32_t *d_buf, long long n, long long first, unsigned long long *d_sum);
int launch_fill(Ctx *c, void *d, size_t bytes);

Ctx *default_ctx(int *status);   // per-thread context behind the reference-signature one-shot entry points

// ---- tuned kernels (kernels_fast.cu) ------------------------------------------------------------------------------
// complex data, real taps, decimate by D.  Returns SDR_OK and sets *done to the number of leading outputs it
// produced (a multiple of its tile; 0 when the shape / alignment has no tuned kernel); the caller finishes the rest
// with launch_fir_generic.  *name receives a static string naming the instantiation.
// Padded-segment ring (kernels_fast.cu): real or complex data, decimation 4 / 8 / 16, up to 128 stored taps.
// x = seg.a ++ seg.b.  When both sources and their boundary are 16-byte aligned the kernel covers ALL `num` outputs
// (windows straddling the two segments and the ragged last tile included): *done == num; otherwise *done = the leading
// outputs it produced (a multiple of its tile; 0: no tuned kernel for the shape) and the caller finishes the rest with
// launch_fir_generic.  d_taps must be zero-padded to at least 128 floats (FirRec does that).
int launch_dec_fast(Ctx *c, bool cplx, int taps_stored, int D, const float *d_taps, Seg2 seg, void *d_out, long long num,
                    long long *done, const char **name, const float *h_taps = nullptr);
// persistent consumer (kernels_fast.cu): see PersistCtl there; ctl = page-locked mapped control block
bool dec_persist_geometry(int taps_stored, int D, bool cplx, int *run_samples, int *halo_samples);
int launch_dec_persist(Ctx *c, int taps_stored, int D, bool cplx, const float *d_taps, const void *d_in, void *d_out, void *ctl,
                       void *d_relay, long long runs_total, cudaStream_t stream, int *grid_out, const char **name);
bool dec_fast_will_cover(bool cplx, int taps_stored, int D, Seg2 seg, long long num);
// opt a kernel in to `smem_bytes` of dynamic shared memory, once per (kernel, device)
int ring_attr(Ctx *c, const void *kernel, int smem_bytes);

// contiguous-slot ring (kernels_real.cu): stride 1 (the filters) and 2, real or complex data, up to 128 stored taps
int launch_fir_small_stride_fast(Ctx *c, bool cplx, int taps_stored, int D, const float *d_taps, const void *d_in, long long n_in,
                                 void *d_out, long long num, long long *done, const char **name);
// rational resampler, real or complex data: output 0 is phase 0 and its window starts at d_in; d_plain_taps = the n_taps plain taps
int launch_res_fast(Ctx *c, bool cplx, int L, int M, int n_taps, const float *d_plain_taps, const void *d_in, long long n_in,
                    void *d_out, long long num, long long *done, const char **name);

// fused u8 convert + decimate + FM demod of outputs [0, num) of a byte stream (kernels_fm.cu).  The stream is d_in
// (a_samples IQ pairs) followed by d_in_b, n_samples pairs in all (a_samples == n_samples: one segment).
int launch_fm_front(Ctx *c, int T, int D, const float *d_taps, bool symmetric, const uint8_t *d_in, long long a_samples,
                    const uint8_t *d_in_b, long long n_samples, float *d_out, long long num, float2 *d_bnd,
                    long long bnd_capacity_subtiles, const float2 *d_carry, float2 *d_carry_out, unsigned int *d_ticket,
                    long long *done, const char **name);
// fused u8 convert + decimate (complex outputs), same kernel without the discriminator
int launch_dec_u8(Ctx *c, int T, int D, const float *d_taps, bool symmetric, const uint8_t *d_in, long long a_samples,
                  const uint8_t *d_in_b, long long n_samples, float *d_out, long long num, long long *done, const char **name);
// firResampler >-> firFilter >-> P.map (* k) fused (kernels_lowrate.cu): outputs z[n], n in [n0, n0 + num), of the flat
// stream z[n] = k * sum_t cf[t] r[n + t], r = the L/M resampling of x (taps cr); x = seg.a ++ seg.b and x[0] is the first
// input sample of resampler output n0 (i.e. global sample ceil(n0 M / L)).  *done = num when a tuned kernel exists.
int launch_fm_lowrate(Ctx *c, int L, int M, int n_taps_r, const float *h_taps_r, int n_taps_f, const float *h_taps_f, float scale,
                      Seg2 seg, long long n0, float *d_out, long long num, long long *done, const char **name);

}  // namespace sdr
