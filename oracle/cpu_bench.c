/*
 * cpu_bench.c -- TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Native driver that times a CPU implementation of the headline path exactly as the reference's firDecimator
 * would issue it (hs_sources/SDR/Filter.hs:578-611) over a stream of equal-sized input vectors:
 *   per vector: decimateOne   count = (len - numCoeffs) / D + 1 outputs          (:588)   <- the C kernel under test
 *               decimateCross count = quotUp(len - count*D, D) outputs           (:604)   <- strict left-to-right sums
 *                                                                                          (FilterInternal.hs:398-402)
 * `one` is a function pointer with the reference's own C signature (decimate.c:105), so the same driver times the
 * compiled UNMODIFIED reference (oracle/_ref/libsdrref.so, kind "reference") or the plain-C port in
 * sdr_oracle.c (kind "port").  Threads process disjoint contiguous ranges of vectors (independent streams), which is
 * how a multi-core host would be used: the reference itself is single-threaded per pipeline.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef void (*decim_fn)(int num, int factor, int numCoeffs, float *coeffs, float *inBuf, float *outBuf);

void o_decimateC(int variant, int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf);
void o_decimateCrossC(int factor, int numCoeffs, const float *coeffs, int num, const float *last, int nlast,
                      const float *next, int nnext, float *out);

/* adapter: the port's AVX-order complex decimator under the reference's signature */
void o_port_decimateAVXRC(int num, int factor, int numCoeffs, float *coeffs, float *inBuf, float *outBuf) {
    o_decimateC(2 /* V_AVX */, num, factor, numCoeffs, coeffs, inBuf, outBuf);
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

struct job {
    decim_fn     one;
    int          factor, n_taps, vec_len;
    const float *dup, *plain, *stream;
    long         v0, v1;   /* vectors [v0, v1) */
    double       t_end;
    long         vectors_done;
    float       *out;
};

static void *worker(void *arg) {
    struct job *j = (struct job *)arg;
    const int   T = j->n_taps, D = j->factor, len = j->vec_len;
    const int   count = (len - T) / D + 1;
    const int   rest = len - count * D;                 /* samples left in the vector: < T            */
    const int   ccount = (rest + D - 1) / D;            /* outputs of the crossover state             */
    long        done = 0;
    do {
        for (long v = j->v0; v < j->v1; v++) {
            const float *in = j->stream + 2 * (size_t)v * len;
            j->one(count, D, 2 * T, (float *)j->dup, (float *)in, j->out);
            if (v + 1 < j->v1 && ccount > 0)
                o_decimateCrossC(D, T, j->plain, ccount, in + 2 * (size_t)count * D, rest, in + 2 * (size_t)len, len,
                                 j->out + 2 * count);
            done++;
        }
    } while (now_s() < j->t_end);
    j->vectors_done = done;
    return NULL;
}

/* returns input samples per second; *samples_done = total input samples consumed, *seconds = wall time */
double o_bench_fir_decimator(void *one, int factor, int n_taps, const float *taps_dup, const float *taps_plain,
                             const float *stream, long n_vectors, int vec_len, int threads, double min_seconds,
                             long *samples_done, double *seconds) {
    if (threads < 1) threads = 1;
    if (threads > n_vectors) threads = (int)n_vectors;
    pthread_t  *th = (pthread_t *)calloc(threads, sizeof(pthread_t));
    struct job *jobs = (struct job *)calloc(threads, sizeof(struct job));
    double      t0 = now_s();
    for (int t = 0; t < threads; t++) {
        jobs[t].one = (decim_fn)one; jobs[t].factor = factor; jobs[t].n_taps = n_taps; jobs[t].vec_len = vec_len;
        jobs[t].dup = taps_dup; jobs[t].plain = taps_plain; jobs[t].stream = stream;
        jobs[t].v0 = n_vectors * t / threads; jobs[t].v1 = n_vectors * (t + 1) / threads;
        jobs[t].t_end = t0 + min_seconds;
        jobs[t].out = (float *)malloc(sizeof(float) * 2 * (vec_len / factor + 64));
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    long total = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); total += jobs[t].vectors_done; free(jobs[t].out); }
    double dt = now_s() - t0;
    free(th); free(jobs);
    if (samples_done) *samples_done = total * vec_len;
    if (seconds) *seconds = dt;
    return (double)total * vec_len / dt;
}

/* ---- u8-fed chain: P.map interleavedIQUnsignedByteToFloat >-> firDecimator (examples/fm/fm.hs:34-36) --------------------
 * per 16384-byte vector: convert (hs_sources/SDR/Util.hs:128-134 -> convert.c:37) into a fresh float vector, then the
 * firDecimator call pattern above on the converted vectors. */
typedef void (*conv_fn)(int num, unsigned char *in, float *out);

struct job8 {
    conv_fn      conv;
    decim_fn     one;
    int          factor, n_taps, vec_len;
    const float *dup, *plain;
    const unsigned char *stream;
    long         v0, v1;
    double       t_end;
    long         vectors_done;
    float       *out, *cur, *nxt;
};

static void *worker8(void *arg) {
    struct job8 *j = (struct job8 *)arg;
    const int    T = j->n_taps, D = j->factor, len = j->vec_len;
    const int    count = (len - T) / D + 1;
    const int    rest = len - count * D;
    const int    ccount = (rest + D - 1) / D;
    long         done = 0;
    do {
        j->conv(2 * len, (unsigned char *)j->stream + 2 * (size_t)j->v0 * len, j->cur);
        for (long v = j->v0; v < j->v1; v++) {
            if (v + 1 < j->v1) j->conv(2 * len, (unsigned char *)j->stream + 2 * (size_t)(v + 1) * len, j->nxt);
            j->one(count, D, 2 * T, (float *)j->dup, j->cur, j->out);
            if (v + 1 < j->v1 && ccount > 0)
                o_decimateCrossC(D, T, j->plain, ccount, j->cur + 2 * (size_t)count * D, rest, j->nxt, len, j->out + 2 * count);
            float *t = j->cur; j->cur = j->nxt; j->nxt = t;
            done++;
        }
    } while (now_s() < j->t_end);
    j->vectors_done = done;
    return NULL;
}

double o_bench_u8_fir_decimator(void *conv, void *one, int factor, int n_taps, const float *taps_dup, const float *taps_plain,
                                const unsigned char *stream, long n_vectors, int vec_len, int threads, double min_seconds,
                                long *samples_done, double *seconds) {
    if (threads < 1) threads = 1;
    if (threads > n_vectors) threads = (int)n_vectors;
    pthread_t   *th = (pthread_t *)calloc(threads, sizeof(pthread_t));
    struct job8 *jobs = (struct job8 *)calloc(threads, sizeof(struct job8));
    double       t0 = now_s();
    for (int t = 0; t < threads; t++) {
        jobs[t].conv = (conv_fn)conv; jobs[t].one = (decim_fn)one; jobs[t].factor = factor; jobs[t].n_taps = n_taps;
        jobs[t].vec_len = vec_len; jobs[t].dup = taps_dup; jobs[t].plain = taps_plain; jobs[t].stream = stream;
        jobs[t].v0 = n_vectors * t / threads; jobs[t].v1 = n_vectors * (t + 1) / threads;
        jobs[t].t_end = t0 + min_seconds;
        jobs[t].out = (float *)malloc(sizeof(float) * 2 * (vec_len / factor + 64));
        jobs[t].cur = (float *)malloc(sizeof(float) * 2 * (vec_len + 64));
        jobs[t].nxt = (float *)malloc(sizeof(float) * 2 * (vec_len + 64));
        pthread_create(&th[t], NULL, worker8, &jobs[t]);
    }
    long total = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL); total += jobs[t].vectors_done;
        free(jobs[t].out); free(jobs[t].cur); free(jobs[t].nxt);
    }
    double dt = now_s() - t0;
    free(th); free(jobs);
    if (samples_done) *samples_done = total * vec_len;
    if (seconds) *seconds = dt;
    return (double)total * vec_len / dt;
}
