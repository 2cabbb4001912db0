// ctx.cu -- error plumbing, device context, memory / event helpers of libsdr_b200.
// Replaces nothing in the reference (it has no device); sdr_has_cuda() is the predicate a `featureSelect` entry
// (reference hs_sources/SDR/CPUID.hs:100-104) would test.
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>

namespace sdr {

static thread_local char g_err[512] = "";

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    const char *base = strrchr(file, '/');
    int code = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? SDR_ENODEVICE
             : (e == cudaErrorMemoryAllocation)                              ? SDR_ENOMEM
                                                                             : SDR_ECUDA;
    cudaGetLastError();  // clear the sticky-less error so later calls report their own
    return set_error(code, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e),
                     base ? base + 1 : file, line, what);
}

int Ctx::bind() const {
    SDR_CUDA(cudaSetDevice(device));
    return SDR_OK;
}

static int grow(void **p, size_t *have, size_t want, bool pinned) {
    if (*have >= want) return SDR_OK;
    size_t cap = *have ? *have : (size_t)1 << 16;
    while (cap < want) cap *= 2;
    if (*p) {
        if (pinned) SDR_CUDA(cudaFreeHost(*p)); else SDR_CUDA(cudaFree(*p));
        *p = nullptr; *have = 0;
    }
    if (pinned) SDR_CUDA(cudaMallocHost(p, cap)); else SDR_CUDA(cudaMalloc(p, cap));
    *have = cap;
    return SDR_OK;
}

int Ctx::ensure_stage(size_t in_bytes, size_t out_bytes) {
    SDR_TRY(grow(&d_stage_in, &d_stage_in_bytes, in_bytes + 256, false));
    SDR_TRY(grow(&d_stage_out, &d_stage_out_bytes, out_bytes + 256, false));
    return SDR_OK;
}

int Ctx::ensure_dc_scratch(size_t bytes) {
    const size_t before = d_dc_scratch_bytes;
    SDR_TRY(grow(&d_dc_scratch, &d_dc_scratch_bytes, bytes, false));
    if (d_dc_scratch_bytes != before) dc_scratch_regrown = true;   // fresh block: the counters in its head are garbage
    return SDR_OK;
}

int ring_attr(Ctx *c, const void *kernel, int smem_bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> seen;
    std::lock_guard<std::mutex> lock(mu);
    if (seen.count({kernel, c->device})) return SDR_OK;
    SDR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    seen.insert({kernel, c->device});
    return SDR_OK;
}

// Per-thread default context used by the reference-signature one-shot entry points (layer 1).
Ctx *default_ctx(int *status) {
    static thread_local Ctx *c = nullptr;
    if (c) { *status = c->bind(); return c; }
    int dev = 0;
    if (const char *e = getenv("SDR_B200_DEVICE")) dev = atoi(e);
    sdr_ctx_t *out = nullptr;
    *status = sdr_ctx_create(dev, &out);
    c = reinterpret_cast<Ctx *>(out);
    return c;
}

}  // namespace sdr

using namespace sdr;

extern "C" {

const char *sdr_last_error(void) { return g_err; }
int sdr_b200_abi_version(void) { return 1; }

int sdr_device_count(int *count) {
    if (!count) return set_error(SDR_EINVAL, "sdr_device_count: null argument");
    *count = 0;
    SDR_CUDA(cudaGetDeviceCount(count));
    if (*count == 0) return set_error(SDR_ENODEVICE, "no CUDA device");
    return SDR_OK;
}

int sdr_has_cuda(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10 && p.minor == 0) return 1;   // sm_100a code only
    }
    return 0;
}

int sdr_ctx_create(int device, sdr_ctx_t **ctx) {
    if (!ctx) return set_error(SDR_EINVAL, "sdr_ctx_create: null argument");
    *ctx = nullptr;
    int n = 0;
    SDR_TRY(sdr_device_count(&n));
    if (device < 0 || device >= n) return set_error(SDR_EINVAL, "sdr_ctx_create: device %d out of range (%d)", device, n);
    cudaDeviceProp p;
    SDR_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10 || p.minor != 0)
        return set_error(SDR_ENODEVICE, "device %d (%s, sm_%d%d) is not sm_100: this library carries sm_100a code only",
                         device, p.name, p.major, p.minor);
    Ctx *c = new Ctx();
    c->device = device;
    c->sm_count = p.multiProcessorCount;
    if (const char *e = getenv("SDR_B200_FIR_FFA")) c->fir_ffa = atoi(e) != 0;
    SDR_CUDA(cudaSetDevice(device));
    SDR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SDR_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    SDR_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    SDR_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    *ctx = reinterpret_cast<sdr_ctx_t *>(c);
    return SDR_OK;
}

int sdr_ctx_destroy(sdr_ctx_t *ctx) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return SDR_OK;
    SDR_TRY(c->bind());
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->side);
    if (c->d_stage_in) cudaFree(c->d_stage_in);
    if (c->d_stage_out) cudaFree(c->d_stage_out);
    if (c->d_flush) cudaFree(c->d_flush);
    if (c->d_dc_scratch) cudaFree(c->d_dc_scratch);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->side);
    delete c;
    return SDR_OK;
}

int sdr_ctx_set_fast_fir(sdr_ctx_t *ctx, int on) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return set_error(SDR_EINVAL, "sdr_ctx_set_fast_fir: null ctx");
    c->fir_ffa = on != 0;
    return SDR_OK;
}

int sdr_ctx_sync(sdr_ctx_t *ctx) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return set_error(SDR_EINVAL, "sdr_ctx_sync: null ctx");
    SDR_TRY(c->bind());
    SDR_CUDA(cudaStreamSynchronize(c->stream));
    SDR_CUDA(cudaStreamSynchronize(c->side));
    return SDR_OK;
}

int sdr_ctx_set_arith(sdr_ctx_t *ctx, int arith_mode) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || (arith_mode != SDR_ARITH_FAST && arith_mode != SDR_ARITH_EXACT))
        return set_error(SDR_EINVAL, "sdr_ctx_set_arith: bad argument");
    c->arith = arith_mode;
    return SDR_OK;
}

int sdr_ctx_sm_count(sdr_ctx_t *ctx, int *sms) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !sms) return set_error(SDR_EINVAL, "sdr_ctx_sm_count: bad argument");
    *sms = c->sm_count;
    return SDR_OK;
}

int sdr_dev_alloc(sdr_ctx_t *ctx, size_t bytes, void **dptr) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !dptr) return set_error(SDR_EINVAL, "sdr_dev_alloc: bad argument");
    SDR_TRY(c->bind());
    SDR_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return SDR_OK;
}

int sdr_dev_free(sdr_ctx_t *ctx, void *dptr) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return set_error(SDR_EINVAL, "sdr_dev_free: null ctx");
    SDR_TRY(c->bind());
    if (dptr) SDR_CUDA(cudaFree(dptr));
    return SDR_OK;
}

int sdr_host_alloc_pinned(size_t bytes, void **hptr) {
    if (!hptr) return set_error(SDR_EINVAL, "sdr_host_alloc_pinned: null argument");
    SDR_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
    return SDR_OK;
}

int sdr_host_free_pinned(void *hptr) {
    if (hptr) SDR_CUDA(cudaFreeHost(hptr));
    return SDR_OK;
}

int sdr_memcpy_h2d(sdr_ctx_t *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return set_error(SDR_EINVAL, "sdr_memcpy_h2d: null ctx");
    SDR_TRY(c->bind());
    if (bytes) SDR_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, c->stream));
    return SDR_OK;
}

int sdr_memcpy_d2h(sdr_ctx_t *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return set_error(SDR_EINVAL, "sdr_memcpy_d2h: null ctx");
    SDR_TRY(c->bind());
    if (bytes) SDR_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    return SDR_OK;
}

int sdr_memcpy_d2d(sdr_ctx_t *ctx, void *dst_dev, const void *src_dev, size_t bytes) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return set_error(SDR_EINVAL, "sdr_memcpy_d2d: null ctx");
    SDR_TRY(c->bind());
    if (bytes) SDR_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return SDR_OK;
}

int sdr_memset_dev(sdr_ctx_t *ctx, void *dst_dev, int byte, size_t bytes) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return set_error(SDR_EINVAL, "sdr_memset_dev: null ctx");
    SDR_TRY(c->bind());
    if (bytes) SDR_CUDA(cudaMemsetAsync(dst_dev, byte, bytes, c->stream));
    return SDR_OK;
}

struct sdr_event { cudaEvent_t ev; int device; };

int sdr_event_create(sdr_ctx_t *ctx, sdr_event_t **ev) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !ev) return set_error(SDR_EINVAL, "sdr_event_create: bad argument");
    SDR_TRY(c->bind());
    sdr_event *e = new sdr_event();
    e->device = c->device;
    cudaError_t r = cudaEventCreate(&e->ev);
    if (r != cudaSuccess) { delete e; return cuda_fail(r, "cudaEventCreate", __FILE__, __LINE__); }
    *ev = e;
    return SDR_OK;
}

int sdr_event_record(sdr_ctx_t *ctx, sdr_event_t *ev) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !ev) return set_error(SDR_EINVAL, "sdr_event_record: bad argument");
    SDR_TRY(c->bind());
    SDR_CUDA(cudaEventRecord(ev->ev, c->stream));
    return SDR_OK;
}

int sdr_event_elapsed_ms(sdr_event_t *start, sdr_event_t *stop, float *ms) {
    if (!start || !stop || !ms) return set_error(SDR_EINVAL, "sdr_event_elapsed_ms: bad argument");
    SDR_CUDA(cudaSetDevice(stop->device));
    SDR_CUDA(cudaEventSynchronize(stop->ev));
    SDR_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
    return SDR_OK;
}

int sdr_event_destroy(sdr_event_t *ev) {
    if (!ev) return SDR_OK;
    cudaSetDevice(ev->device);
    cudaEventDestroy(ev->ev);
    delete ev;
    return SDR_OK;
}

int sdr_ctx_launch_count(sdr_ctx_t *ctx, long long *count) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !count) return set_error(SDR_EINVAL, "sdr_ctx_launch_count: bad argument");
    *count = c->launches;
    return SDR_OK;
}

}  // extern "C"
