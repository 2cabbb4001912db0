/*
 * sdr_b200.h -- C ABI of the B200-native streaming-FIR hot path (drop-in for adamwalker/sdr's native layer).
 *
 * Everything here is `extern "C"`, plain pointers and sizes.  Citations `file:line` are relative to the reference
 * tree (adamwalker/sdr @ 5fcd15c) and name the interface each entry point replaces.
 *
 * Three layers, mirroring the reference's L0-L3:
 *
 *   1. Reference-signature one-shot kernels on HOST pointers            <- the c_sources C symbols imported by
 *      (filterCuda*, decimateCuda*, resampleCuda*, convertCuda*, ...)      FilterInternal.hs:80-249,346-388, Util.hs:100-241
 *   2. Plugin records: sdr_filter / sdr_decimator / sdr_resampler        <- data Filter/Decimator/Resampler, Filter.hs:116-144
 *      with *_one / *_cross closures on host OR device buffers              and their fast-, mk-constructors :163-502
 *   3. Pipes: sdr_pipe_* (push = await, pop = yield)                     <- firFilter/firDecimator/firResampler Filter.hs:532-727,
 *                                                                           fmDemod Demod.hs:38-46, P.map convert/scale (fm.hs:34-40)
 *   plus: device-memory helpers, the multi-GPU shard plan + halo exchange (no reference
 *   counterpart), synthetic-stream generator and measurement helpers used by bench.py.
 *
 * Conventions
 *   - Every function returns int status: 0 = SDR_OK, non-zero = error; sdr_last_error() gives the thread-local
 *     message.  (Layer 1: the reference's C returns void, so `IO ()` becomes `IO CInt`; the one family that already
 *     returns a value, resample*, keeps it -- the next group -- and reports failure as a negative number.)  The reference's C returns void and is UB on bad sizes; its Haskell side raises `error "filter 1"`
 *     etc. on precondition failure (Filter.hs:525-527,544,586,691) -- the same preconditions are checked here
 *     and reported as SDR_EPRECOND with the reference's own location string.
 *   - Complex data is interleaved (re, im) float32, exactly as Data.Complex Float is laid out by
 *     VS.unsafeCast (FilterInternal.hs:73-78).  Counts are in SAMPLES (a complex sample is one sample) unless
 *     a parameter says "bytes"/"floats".
 *   - No entry point retains a caller pointer after it returns (FilterInternal.hs:68-71 pins only for the call).
 *   - Any OS thread may call any function; a handle must not be used by two threads at once (the reference is
 *     single-threaded per pipeline).
 *   - There is NO CPU fallback: with no usable CUDA device every compute entry point fails with SDR_ENODEVICE.
 */
#ifndef SDR_B200_H
#define SDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------------------- */
/* status                                                                                                     */
/* ---------------------------------------------------------------------------------------------------------- */
enum {
    SDR_OK        = 0,
    SDR_EINVAL    = 1, /* bad argument                                                         */
    SDR_EPRECOND  = 2, /* a reference `assert` would have fired (Filter.hs:525-527)            */
    SDR_ECUDA     = 3, /* CUDA runtime error (message carries cudaGetErrorString)              */
    SDR_ENODEVICE = 4, /* no CUDA device / driver: the product has no CPU path                 */
    SDR_ENOMEM    = 5,
    SDR_ENCCL     = 6, /* NCCL missing or failed                                               */
    SDR_EAGAIN    = 7  /* sdr_pipe_pop: no complete output block yet                           */
};

/* where a buffer lives */
enum {
    SDR_HOST = 0,        /* any host memory; the call returns only when the library no longer needs the pointer */
    SDR_DEVICE = 1,      /* device memory of the handle's context; enqueue-only on the ctx stream             */
    SDR_HOST_PINNED = 2, /* sdr_pipe_push / sdr_pipe_pop only: page-locked host memory the caller leaves untouched
                            until sdr_pipe_sync -- enqueue-only, no staging copy, copies of adjacent vectors merge */
    SDR_DEVICE_HELD = 3  /* sdr_pipe_push / sdr_pipe_run only: device memory the caller leaves untouched until sdr_pipe_sync
                            -- ZERO-COPY: FIR stages read the vector in place (vectors adjacent in memory form one run that
                            the next launch reads together with the stage's carried tail); only the few samples a launch
                            leaves over are copied into the stage.  This is the Pipes zero-copy contract on the device:
                            a yielded vector is immutable and the stage may keep referring to it (Filter.hs:519-521) */
};

/* arithmetic mode of the FIR kernels */
enum {
    SDR_ARITH_FAST  = 0, /* fused multiply-add, taps summed in increasing order (default; <= 1e-5 of output scale) */
    SDR_ARITH_EXACT = 1  /* unfused mul+add in the reference's AVX lane order (common.h:18-29,58-72,82-90):
                            bit-identical to the *AVX* C variants; test / verification mode, ~3x slower        */
};

const char *sdr_last_error(void);
int         sdr_b200_abi_version(void); /* bumps when this header changes incompatibly */

/* ---------------------------------------------------------------------------------------------------------- */
/* context: one device + one CUDA stream + pinned staging.  Replaces nothing in the reference (it has no       */
/* device); plays the role CPUInfo/getCPUInfo play for dispatch (CPUID.hs:61-75).                              */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sdr_ctx sdr_ctx_t;

int sdr_device_count(int *count);                       /* SDR_ENODEVICE when the driver is absent */
int sdr_ctx_create(int device, sdr_ctx_t **ctx);
int sdr_ctx_destroy(sdr_ctx_t *ctx);
int sdr_ctx_sync(sdr_ctx_t *ctx);                       /* wait for everything enqueued on the ctx stream */
int sdr_ctx_set_arith(sdr_ctx_t *ctx, int arith_mode);  /* SDR_ARITH_* for handles created afterwards */
/* Real stride-1 filters of 32 / 64 / 128 taps (fastFilterR / fastFilterSymR, Filter.hs:193-261): evaluate the tuned
 * kernel with 2-parallel fast-FIR arithmetic (3 half-length sub-filters per 2 outputs, 17 % fewer FP32 operations;
 * sdr_b200/csrc/fir_ffa.cuh).  Agrees with the direct form to rounding (1.5e-6 of the output scale, bar 1e-5), not
 * bit for bit; off by default (environment default: SDR_B200_FIR_FFA=1). */
int sdr_ctx_set_fast_fir(sdr_ctx_t *ctx, int on);
/* `hasCUDA` predicate for featureSelect (CPUID.hs:100-104): 1 when a sm_100 device is usable, else 0 */
int sdr_has_cuda(void);
int sdr_ctx_sm_count(sdr_ctx_t *ctx, int *sms);

/* device / pinned memory helpers (tests, bench and foreign bindings use these; no torch types anywhere) */
int sdr_dev_alloc(sdr_ctx_t *ctx, size_t bytes, void **dptr);
int sdr_dev_free(sdr_ctx_t *ctx, void *dptr);
int sdr_host_alloc_pinned(size_t bytes, void **hptr);
int sdr_host_free_pinned(void *hptr);
int sdr_memcpy_h2d(sdr_ctx_t *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async on ctx stream */
int sdr_memcpy_d2h(sdr_ctx_t *ctx, void *dst_host, const void *src_dev, size_t bytes); /* async on ctx stream */
int sdr_memcpy_d2d(sdr_ctx_t *ctx, void *dst_dev, const void *src_dev, size_t bytes); /* async on ctx stream */
int sdr_memset_dev(sdr_ctx_t *ctx, void *dst_dev, int byte, size_t bytes);

/* timing on the ctx stream (bench.py): CUDA events */
typedef struct sdr_event sdr_event_t;
int sdr_event_create(sdr_ctx_t *ctx, sdr_event_t **ev);
int sdr_event_record(sdr_ctx_t *ctx, sdr_event_t *ev);
int sdr_event_elapsed_ms(sdr_event_t *start, sdr_event_t *stop, float *ms); /* syncs on `stop` */
int sdr_event_destroy(sdr_event_t *ev);
/* number of kernels this library launched on this ctx since creation (bench.py's gpu_launches) */
int sdr_ctx_launch_count(sdr_ctx_t *ctx, long long *count);

/* ---------------------------------------------------------------------------------------------------------- */
/* Layer 1: reference-signature one-shot kernels, HOST pointers in and out.                                    */
/* Same argument order and meaning as the C symbol each replaces; `num` = number of OUTPUTS.  They stage        */
/* through an internal per-thread context on device 0 (or $SDR_B200_DEVICE) and block until `out` is written.   */
/* ---------------------------------------------------------------------------------------------------------- */

/* filterRR/filterSSERR/filterAVXRR (filter.c:16,27,37): coeffs = numCoeffs taps */
int filterCudaRR(int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);
/* filterSSESymmetricRR/filterAVXSymmetricRR (filter.c:50,60): coeffs = FIRST HALF, numCoeffs = half length */
int filterCudaSymmetricRR(int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);
/* filterRC/filterSSERC2/filterAVXRC2 (filter.c:74,96,116): complex data, numCoeffs plain real taps */
int filterCudaRC(int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);
/* filterSSERC/filterAVXRC (filter.c:86,106): coeffs holds each tap TWICE, numCoeffs = 2*taps (Filter.hs:206) */
int filterCudaRCDup(int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);
/* filterSSESymmetricRC/filterAVXSymmetricRC (filter.c:129,139) */
int filterCudaSymmetricRC(int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);

/* decimate* (decimate.c:16-146), same five coefficient layouts */
int decimateCudaRR(int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);
int decimateCudaSymmetricRR(int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);
int decimateCudaRC(int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);
int decimateCudaRCDup(int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);
int decimateCudaSymmetricRC(int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);

/* resample2RR/resampleSSERR/resampleAVXRR and resample2RC/resampleSSERC/resampleAVXRC (resample.c:34-142): the
 * reference's EXACT signature -- polyphase group tables as FilterInternal.mkResampler builds them (:335-342), and the
 * RETURN VALUE is the next group index (>= 0), which mkResampler's binding (:358-362) threads into the next call.
 * Failure: -(SDR_E*) < 0, message in sdr_last_error(). */
int resampleCudaRR(int buf_size, int num_coeffs, int starting_group, int num_groups, int *increments, float **coeffs,
                   float *in_buf, float *out_buf);
int resampleCudaRC(int buf_size, int num_coeffs, int starting_group, int num_groups, int *increments, float **coeffs,
                   float *in_buf, float *out_buf);
/* resampleRR (resample.c:16-32), legacy single-array form */
int resampleCudaLegacyRR(int buf_size, int coeff_size, int interpolation, int decimation, int filter_offset,
                         const float *coeffs, const float *in_buf, float *out_buf);

/* convertC/convertCSSE/convertCAVX (convert.c:15,22,37): num = BYTES (Util.hs:133); bit-exact */
int convertCuda(int num, const uint8_t *in, float *out);
/* convertCBladeRF/SSE/AVX (convert.c:52,59,73): num = int16 components; bit-exact */
int convertCudaBladeRF(int num, const int16_t *in, float *out);
/* convertBladeRFTransmit (convert.c:87): bit-exact */
int convertCudaBladeRFTransmit(int num, const float *in, int16_t *out);
/* scale/scaleSSE/scaleAVX (scale.c:15,22,30): bit-exact (one rounding) */
int scaleCuda(int num, float factor, const float *in_buf, float *out_buf);
/* dcBlocker (filter.c:152): y[n] = x[n] - x[n-1] + 0.997*y[n-1], product in double */
int dcBlockerCuda(int num, float lastSample, float lastOutput, float *finalSample, float *finalOutput,
                  const float *inBuf, float *outBuf);
/* fmDemodVec (Demod.hs:32-36; no C in the reference): num complex samples in, num floats out; (lastRe,lastIm) is
 * the previous buffer's final sample, (0,0) at stream start (Demod.hs:41) */
int fmDemodCuda(int num, float lastRe, float lastIm, const float *in, float *out);

/* the element-wise kernels on DEVICE-resident buffers (enqueue-only on the ctx stream) */
int sdr_dev_convert_u8(sdr_ctx_t *ctx, const uint8_t *d_in, float *d_out, long long n_bytes);   /* convert.c:15 */
int sdr_dev_scale(sdr_ctx_t *ctx, float factor, const float *d_in, float *d_out, long long n);  /* scale.c:15   */
int sdr_dev_fm_demod(sdr_ctx_t *ctx, float last_re, float last_im, const float *d_in, float *d_out, long long n); /* Demod.hs:32 */
/* dcBlocker (filter.c:152) on device buffers; d_final2 receives (finalSample, finalOutput).  Vectors of at least
 * 65536 samples are evaluated chunk-parallel by speculation + verification (sdr_b200/csrc/dc_spec.cuh) -- still
 * bit-exact for any input; shorter, unaligned or in-place calls run the serial one-lane kernel. */
int sdr_dev_dc_blocker(sdr_ctx_t *ctx, float last_sample, float last_output, const float *d_in, float *d_out, long long n,
                       float *d_final2);
/* Tuning of that speculation for this context: chunk length in samples (0 = automatic), cheap and exact warm-up
 * lengths (-1 = default 6144 / 4096), shortest vector that takes the parallel path (-1 = default).  Results are
 * bit-exact for every setting; short warm-ups only make chunks miss and be repaired serially. */
int sdr_dc_blocker_tuning(sdr_ctx_t *ctx, int chunk, int cheap_warmup, int exact_warmup, long long min_parallel);
/* Counters since the context's first parallel call: stats[0] parallel calls, [1] chunks, [2] chunks repaired,
 * [3] samples rewritten by repairs; *last_parallel = whether the most recent dcBlocker call took the parallel path.
 * Synchronises the ctx stream. */
int sdr_dc_blocker_stats(sdr_ctx_t *ctx, long long stats[4], int *last_parallel);

/* ---------------------------------------------------------------------------------------------------------- */
/* Verification entry points: SDR_ARITH_EXACT arithmetic of ONE named reference variant, HOST pointers.           */
/* `coeffs` / `numCoeffs` exactly as that reference function receives them (plain, duplicated or half).           */
/* Bit-identical to the reference C built with its own flags (sdr.cabal:114, no FMA contraction).                 */
/* ---------------------------------------------------------------------------------------------------------- */
enum {
    SDR_V_SCALAR = 0, /* filterRR / filterRC / decimateRR / decimateRC / resample2RR / resample2RC (common.h:34,95)   */
    SDR_V_SSE    = 1, /* the SSERR and SSERC (duplicated coefficients) functions (common.h:43)                        */
    SDR_V_AVX    = 2, /* the AVXRR and AVXRC (duplicated coefficients) functions (common.h:58)                        */
    SDR_V_SSE2   = 3, /* SSERC2, resampleSSERC (common.h:107)                                                          */
    SDR_V_AVX2   = 4, /* AVXRC2, resampleAVXRC (common.h:129)                                                          */
    SDR_V_SSESYM = 5, /* SSESymmetricRR / RC (common.h:160,206)                                                        */
    SDR_V_AVXSYM = 6  /* AVXSymmetricRR / RC (common.h:181,235)                                                        */
};
/* filter.c / decimate.c families; factor = 1 for the filters */
int sdr_exact_decimate(int variant, int is_complex, int num, int factor, int numCoeffs, const float *coeffs,
                       const float *inBuf, float *outBuf);
/* resample.c:34-142 families; table = num_groups rows of row_stride floats (the reference passes float**) */
int sdr_exact_resample(int variant, int is_complex, int buf_size, int num_coeffs, int starting_group, int num_groups,
                       const int *increments, const float *table, int row_stride, const float *in_buf, float *out_buf,
                       int *next_group);

/* ---------------------------------------------------------------------------------------------------------- */
/* Layer 2: plugin records (Filter.hs:116-144).  Constructors take the taps as the user passes them to the      */
/* reference's fast* constructors.  `mem` says where in/out live (SDR_HOST: staged; SDR_DEVICE: zero-copy,      */
/* enqueue-only on the ctx stream -- call sdr_ctx_sync or use events before reading results).                   */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sdr_filter    sdr_filter_t;
typedef struct sdr_decimator sdr_decimator_t;
typedef struct sdr_resampler sdr_resampler_t;

/* fastFilterR / fastFilterC (Filter.hs:193-196,229-232).  size_multiple pads the stored tap count like mkFilter's
 * roundUp (Filter.hs:169): pass 1 for none (CUDA needs none), 8/4 to reproduce the AVX constructors' numCoeffsF. */
int sdr_filter_create(sdr_ctx_t *ctx, int is_complex, const float *coeffs, int num_coeffs, int size_multiple,
                      sdr_filter_t **f);
/* fastFilterSymR (Filter.hs:258-261): coeffs = first half; numCoeffsF = 2*half (Filter.hs:244) */
int sdr_filter_create_sym(sdr_ctx_t *ctx, int is_complex, const float *half_coeffs, int half_len, sdr_filter_t **f);
int sdr_filter_destroy(sdr_filter_t *f);
int sdr_filter_num_coeffs(const sdr_filter_t *f); /* numCoeffsF */
const char *sdr_filter_last_kernel(const sdr_filter_t *f); /* kernel the last call on this handle dispatched to */
/* filterOne  :: Int -> v a -> vm a -> m ()            (Filter.hs:118) */
int sdr_filter_one(sdr_filter_t *f, int count, const void *in, void *out, int mem);
/* filterCross :: Int -> v a -> v a -> vm a -> m ()    (Filter.hs:119; FilterInternal.hs:405-408) */
int sdr_filter_cross(sdr_filter_t *f, int count, const void *last, int n_last, const void *next, int n_next,
                     void *out, int mem);

/* whole device-resident stream in one call: y[n], n < num, over n_in resident samples (enqueue-only) */
int sdr_filter_stream(sdr_filter_t *f, const void *d_in, long long n_in, void *d_out, long long num);

/* fastDecimatorR / fastDecimatorC / fastDecimatorSymR (Filter.hs:311-315,352-356,385-389) */
int sdr_decimator_create(sdr_ctx_t *ctx, int is_complex, int factor, const float *coeffs, int num_coeffs,
                         int size_multiple, sdr_decimator_t **d);
int sdr_decimator_create_sym(sdr_ctx_t *ctx, int is_complex, int factor, const float *half_coeffs, int half_len,
                             sdr_decimator_t **d);
int sdr_decimator_destroy(sdr_decimator_t *d);
int sdr_decimator_num_coeffs(const sdr_decimator_t *d); /* numCoeffsD  */
int sdr_decimator_factor(const sdr_decimator_t *d);     /* decimationD */
/* decimateOne / decimateCross (Filter.hs:129-130; FilterInternal.hs:398-402) */
int sdr_decimate_one(sdr_decimator_t *d, int count, const void *in, void *out, int mem);
int sdr_decimate_cross(sdr_decimator_t *d, int count, const void *last, int n_last, const void *next, int n_next,
                       void *out, int mem);
/* whole device-resident stream in one call: y[m], m < num, over n_in resident samples (enqueue-only); what bench.py
 * times and what the interior of a multi-GPU shard runs */
int sdr_decimate_stream(sdr_decimator_t *d, const void *d_in, long long n_in, void *d_out, long long num);
/* name of the kernel the last sdr_decimate_one on this handle dispatched to ("dec_c_fast<...>", "fir_generic", ...) */
const char *sdr_decimator_last_kernel(const sdr_decimator_t *d);

/* fastResamplerR / fastResamplerC (Filter.hs:468-473,497-502).  The existential `dat` (Filter.hs:141) is the pair
 * (group, offset) (Filter.hs:424): callers thread it through resample_one / resample_cross. */
typedef struct { int group; int offset; } sdr_resampler_dat_t;
int sdr_resampler_create(sdr_ctx_t *ctx, int is_complex, int interpolation, int decimation, const float *coeffs,
                         int num_coeffs, int size_multiple, sdr_resampler_t **r);
int sdr_resampler_destroy(sdr_resampler_t *r);
int sdr_resampler_num_coeffs(const sdr_resampler_t *r);    /* numCoeffsR (Filter.hs:422) */
const char *sdr_resampler_last_kernel(const sdr_resampler_t *r);
int sdr_resampler_interpolation(const sdr_resampler_t *r); /* interpolationR */
int sdr_resampler_decimation(const sdr_resampler_t *r);    /* decimationR */
/* resampleOne :: dat -> Int -> v a -> vm a -> m (dat, Int) (Filter.hs:142); *end_offset = the Int */
int sdr_resample_one(sdr_resampler_t *r, sdr_resampler_dat_t *dat, int count, const void *in, void *out, int mem,
                     int *end_offset);
/* whole device-resident stream in one call, from output 0 of the stream (group 0): enqueue-only */
int sdr_resample_stream(sdr_resampler_t *r, const void *d_in, long long n_in, void *d_out, long long num);
/* resampleCross (Filter.hs:143; FilterInternal.hs:411-423) */
int sdr_resample_cross(sdr_resampler_t *r, sdr_resampler_dat_t *dat, int count, const void *last, int n_last,
                       const void *next, int n_next, void *out, int mem, int *end_offset);

/* ---------------------------------------------------------------------------------------------------------- */
/* Layer 3: Pipes.  A pipe consumes whole input vectors (push = the Pipe's `await`) and produces output vectors   */
/* of exactly block_size_out elements (pop = `yield`), re-blocking like advanceOutBuf (Filter.hs:516-523).        */
/* On the device the stream is kept contiguous (tail carried in HBM), so there is no crossover case.              */
/* FIR stages launch lazily: only when the new outputs complete at least one output vector (sdr_pipe_set_batch     */
/* raises that threshold); since vectors are only ever yielded whole this is invisible except in timing.          */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sdr_pipe sdr_pipe_t;

int sdr_pipe_fir_filter(sdr_filter_t *f, int block_size_out, sdr_pipe_t **p);       /* firFilter    Filter.hs:532 */
int sdr_pipe_fir_decimator(sdr_decimator_t *d, int block_size_out, sdr_pipe_t **p); /* firDecimator Filter.hs:574 */
int sdr_pipe_fir_resampler(sdr_resampler_t *r, int block_size_out, sdr_pipe_t **p); /* firResampler Filter.hs:679 */
int sdr_pipe_fm_demod(sdr_ctx_t *ctx, sdr_pipe_t **p);                              /* fmDemod      Demod.hs:40   */
/* `P.map interleavedIQUnsignedByteToFloat >-> firDecimator d block_size_out >-> fmDemod` (fm.hs:34-37) as ONE fused
 * stage: pushes are u8 IQ bytes (n = byte count, even), yields are vectors of block_size_out float phases.  Same
 * stream, bit for bit, as the three stages connected one after the other. */
int sdr_pipe_fm_frontend(sdr_decimator_t *d, int block_size_out, sdr_pipe_t **p);
/* `P.map interleavedIQUnsignedByteToFloat >-> firDecimator d block_size_out` (fm.hs:34-36) as ONE fused stage: pushes are
 * u8 IQ bytes (n = byte count, even), yields are vectors of block_size_out COMPLEX samples.  Same stream, bit for bit, as
 * the two stages connected one after the other; the host link carries 2 B per input sample instead of 8. */
int sdr_pipe_u8_decimator(sdr_decimator_t *d, int block_size_out, sdr_pipe_t **p);
/* `firResampler r block_size_resampler >-> firFilter f block_size_out >-> P.map (VG.map (* scale))` (fm.hs:38-40) as ONE
 * fused stage: real vectors in, vectors of block_size_out scaled filter outputs out; the resampled stream never touches
 * HBM.  Same stream, bit for bit, as the three stages connected one after the other (the filter only ever sees whole
 * block_size_resampler vectors of resampled samples, as it does behind the un-fused resampler stage). */
int sdr_pipe_fm_lowrate(sdr_resampler_t *r, int block_size_resampler, sdr_filter_t *f, int block_size_out, float scale, sdr_pipe_t **p);
const char *sdr_pipe_last_kernel(const sdr_pipe_t *p);
int sdr_pipe_convert_u8(sdr_ctx_t *ctx, sdr_pipe_t **p); /* P.map interleavedIQUnsignedByteToFloat (Util.hs:104) */
int sdr_pipe_scale(sdr_ctx_t *ctx, float factor, sdr_pipe_t **p);                   /* P.map (VG.map (* k)) fm.hs:40 */
int sdr_pipe_dc_blocker(sdr_ctx_t *ctx, sdr_pipe_t **p);                            /* dcBlockingFilter Filter.hs:730 */
int sdr_pipe_destroy(sdr_pipe_t *p);
/* feed one upstream vector of n INPUT elements (for convert_u8: n bytes).  Fails with SDR_EPRECOND
 * ("filter 1" / "decimate 1" / "resample 1") when the very first / a post-drain vector is shorter than numCoeffs. */
int sdr_pipe_push(sdr_pipe_t *p, const void *in, long long n, int mem);
/* number of complete output vectors ready to pop */
int sdr_pipe_ready(sdr_pipe_t *p, int *n_blocks);
/* length in elements of the vector the next sdr_pipe_pop will deliver (SDR_EAGAIN when none is ready) */
int sdr_pipe_next_len(sdr_pipe_t *p, long long *n);
/* copy out the next output vector (`out` must hold sdr_pipe_next_len elements); elementwise pipes (fm_demod/convert/scale) yield one vector per pushed
 * vector with that vector's length (returned in *n_out).  SDR_EAGAIN when none is ready. */
int sdr_pipe_pop(sdr_pipe_t *p, void *out, long long *n_out, int mem);
/* issue any deferred SDR_HOST_PINNED copies and wait for everything enqueued on the pipe's stream: after this the
 * caller may reuse the pinned vectors it pushed and read pinned / device vectors it popped */
int sdr_pipe_sync(sdr_pipe_t *p);
/* throughput knob for FIR stages: do not launch before `min_outputs` new outputs are computable (default 0 = as soon
 * as one output vector can be completed, the lowest latency).  Yielded vectors are unchanged, only their timing. */
int sdr_pipe_set_batch(sdr_pipe_t *p, long long min_outputs);
/* Per-vector device pushes without a launch per vector (the Pipes per-8192 contract of firDecimator, Filter.hs:578-598, on
 * the device): with max_session_samples > 0, SDR_DEVICE_HELD vectors pushed back to back (adjacent in memory) are consumed by
 * a RESIDENT kernel -- a push is one store into a page-locked control block the kernel polls, sdr_pipe_ready / _pop see the
 * runs it has published as finished.  A session ends at sdr_pipe_sync, at a vector that is not adjacent, or after
 * max_session_samples (the output FIFO is sized for that up front); the ordinary launch path finishes what a session leaves
 * (the carried tail, an incomplete run).  The kernel leaves 7 SMs free; after 3 s without input it winds the session down
 * without hanging the GPU (the next push opens a new session).  Complex decimate-by-8 stages with up to 128 stored taps; 0 switches it off. */
int sdr_pipe_set_persistent(sdr_pipe_t *p, long long max_session_samples);
/* `runEffect $ each vectors >-> p >-> ... >-> sink >-> collect` as one native loop: pushes n_vecs consecutive vectors of
 * vec_len input elements starting at `in` into `p`, pops every vector `sink` yields (sink = p, or the last stage
 * connected behind it) into `out` back to back, then sdr_pipe_sync.  *n_out = elements written (<= out_capacity).
 * With out_mem == SDR_DEVICE a FIR-kind sink produces straight into `out` (no copy of the yielded vectors): the elements
 * between *n_out and out_capacity are then scratch (they may hold the partial output block that the stage keeps for the
 * next call).  Nothing is ever written past out_capacity. */
int sdr_pipe_run(sdr_pipe_t *p, sdr_pipe_t *sink, const void *in, long long vec_len, long long n_vecs, int in_mem,
                 void *out, long long out_capacity, int out_mem, long long *n_out);
/* Producer / consumer edges on file descriptors, the steps either side of the path:
 *   SDR.Serialize.fromHandle n h (Serialize.hs:82-83)  -- stream mode: every vector is exactly vec_len elements read from
 *       in_fd (blocking until complete; a shorter vector only at end of file), until end of file or max_vecs (0 = no limit);
 *   SDR.NetworkStream.udpSource sock size (NetworkStream.hs:28-35) -- SDR_IO_DATAGRAM_IN: one read() = one vector of up
 *       to vec_len elements, max_vecs required;
 *   SDR.Serialize.toHandle (Serialize.hs:78-79) -- every vector `sink` yields is written to out_fd as raw elements
 *       (out_fd < 0: discarded); SDR_IO_DATAGRAM_OUT writes one vector per write() (udpSink, NetworkStream.hs:37-42).
 * read() lands directly in a page-locked ring (two halves of ~8 MiB: one is filled while the other is copied to the
 * device), vectors are pushed as SDR_HOST_PINNED.  A short last vector that violates a FIR stage's minimum length
 * returns SDR_EPRECOND after everything before it has been processed and written (the reference asserts there). */
enum { SDR_IO_DATAGRAM_IN = 1, SDR_IO_DATAGRAM_OUT = 2 };
typedef struct {
    long long vectors_in, elements_in, vectors_out, elements_out;
    double    read_seconds, write_seconds;   /* time spent blocked in read() / write() */
} sdr_io_stats_t;
int sdr_pipe_run_fd(sdr_pipe_t *p, sdr_pipe_t *sink, int in_fd, long long vec_len, long long max_vecs, int out_fd, int flags,
                    sdr_io_stats_t *stats);
/* Stream state export / import (checkpoint / resume).  The state of a stage is what it carries between vectors: the
 * not-yet-consumed tail of the input stream (the reference's crossover carry, Filter.hs:558-569, 600-611, 712-727), the
 * resampler's counters ((group, offset), Filter.hs:419-424), fmDemod's last sample (Demod.hs:41-46), the dcBlocker pair
 * (Filter.hs:731-739) and the outputs not yet popped (advanceOutBuf's partial block, Filter.hs:516-523).  save writes a
 * self-describing blob to HOST memory (sdr_pipe_state_size bytes; synchronises the stage's stream); restore loads it
 * into a stage constructed the same way (kind, taps, factors, block size -- checked, SDR_EINVAL otherwise).  The stream
 * then continues bit for bit as if it had never been interrupted.  Connections are not part of the state. */
int sdr_pipe_state_size(sdr_pipe_t *p, size_t *bytes);
int sdr_pipe_state_save(sdr_pipe_t *p, void *buf, size_t capacity, size_t *written);
int sdr_pipe_state_restore(sdr_pipe_t *p, const void *buf, size_t bytes);
/* connect: everything `src` yields is pushed into `dst` device-to-device without touching the host (>->) */
int sdr_pipe_connect(sdr_pipe_t *src, sdr_pipe_t *dst);

/* ---------------------------------------------------------------------------------------------------------- */
/* Multi-GPU: shard a flat stream over `world` ranks in overlapping chunks.                */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    long long in_begin;   /* first input sample this rank keeps resident                       */
    long long in_count;   /* samples resident on this rank (without halo)                      */
    long long halo;       /* samples needed from rank+1's chunk start (0 on the last rank)     */
    long long out_begin;  /* first global output index this rank owns                          */
    long long out_count;  /* outputs this rank owns                                            */
    long long out_interior; /* of those, how many need no halo (computed before the exchange)  */
    long long n_samples;  /* the arguments the plan was made from (the exchange needs the     */
    int taps, factor, world, rank; /* neighbour's plan too)                                    */
} sdr_shard_t;
/* pure host arithmetic (no device): plan for a decimating FIR with `taps`, `factor` over n_samples */
int sdr_shard_plan(long long n_samples, int taps, int factor, int world, int rank, sdr_shard_t *plan);

typedef struct sdr_comm sdr_comm_t;
/* NCCL bootstrap: rank 0 calls sdr_comm_unique_id, the id bytes travel by any host channel (bench.py uses
 * torch.distributed/gloo purely as that channel), every rank calls sdr_comm_create */
#define SDR_COMM_ID_BYTES 128
int sdr_comm_unique_id(unsigned char id[SDR_COMM_ID_BYTES]);
int sdr_comm_create(sdr_ctx_t *ctx, const unsigned char id[SDR_COMM_ID_BYTES], int world, int rank, sdr_comm_t **c);
int sdr_comm_destroy(sdr_comm_t *c);
/* in-stream rendezvous of all ranks: a 4-byte ncclAllReduce enqueued on the ctx stream (enqueue-only) */
int sdr_comm_barrier(sdr_comm_t *c);
/* Optional peer-memory transport for the halo (collective over the communicator): every rank passes the base address
 * of the sdr_dev_alloc allocation holding its chunk; CUDA IPC handles are exchanged with ncclAllGather and each rank
 * maps its right neighbour's chunk.  sdr_decimate_sharded on that same d_in is then ONE launch: the ring kernel's edge
 * fills read the T-D halo samples in place from the neighbour's HBM over NVLink (TMA bulk copies of peer memory; no
 * per-pass NCCL kernel, no rendezvous).  The caller guarantees the neighbour's chunk is complete before a pass starts. */
int sdr_comm_share_chunks(sdr_comm_t *c, const void *d_chunk_base); /* NULL: drop the mapping (local call), NCCL transport again */
int sdr_comm_peer_halo_active(const sdr_comm_t *c, const void *d_in);
/* one pass of the sharded decimator: interior outputs on the ctx stream, halo exchange (ncclSend of my first
 * `halo` samples to rank-1 / ncclRecv from rank+1) on a side stream, boundary outputs after the halo lands.
 * d_in: this rank's resident chunk (plan.in_count samples); d_out: plan.out_count outputs. */
int sdr_decimate_sharded(sdr_decimator_t *d, sdr_comm_t *c, const sdr_shard_t *plan, const void *d_in, void *d_out);

/* ---------------------------------------------------------------------------------------------------------- */
/* Synthetic streams + measurement helpers (bench.py, tests)                                                     */
/* ---------------------------------------------------------------------------------------------------------- */
/* Counter-based white noise keyed on the GLOBAL float index (so any shard regenerates identical data): float j of
 * the stream = ((sum of four 16-bit fields of two 32-bit hashes of (seed, first_float + j)) - 131070) / 37837.2...
 * Exactly reproducible on the CPU (tests/synth.py).  n_floats floats written at d_out. */
int sdr_synth_noise(sdr_ctx_t *ctx, float *d_out, long long n_floats, long long first_float, uint32_t seed);
/* uniform bytes keyed on the global byte index */
int sdr_synth_bytes(sdr_ctx_t *ctx, uint8_t *d_out, long long n_bytes, long long first_byte, uint32_t seed);
/* writes `bytes` of device scratch to evict L2 between timed iterations */
int sdr_flush_l2(sdr_ctx_t *ctx);
/* position-weighted wrap-around checksum of n_words 32-bit words on the device:
 * sum_i (w[i] + 1) * (2 * (first_word + i) + 1) mod 2^64 -- additive over shards, so sharded == single-GPU can be
 * checked at full size without moving the data; blocks until the sum is on the host */
int sdr_checksum32(sdr_ctx_t *ctx, const void *d_buf, long long n_words, long long first_word, uint64_t *sum);

#ifdef __cplusplus
}
#endif
#endif /* SDR_B200_H */
