// pipes.cu -- layer 3 of the C ABI: the streaming stages of the reference as push/pop handles.
//
//   firFilter / firDecimator / firResampler   hs_sources/SDR/Filter.hs:532-727
//   fmDemod                                   hs_sources/SDR/Demod.hs:38-46
//   P.map interleavedIQUnsignedByteToFloat    hs_sources/SDR/Util.hs:104, examples/fm/fm.hs:35
//   P.map (VG.map (* k))                      examples/fm/fm.hs:40
//
// The reference keeps the stream as separate host vectors and therefore needs a `crossover` state whenever a
// window straddles two of them (Filter.hs:558-569, 600-611, 712-727).  Here the stream is kept CONTIGUOUS in HBM: a
// linear device buffer holds the not-yet-consumed tail followed by every newly pushed vector, so each push is the
// closed-form flat-stream computation over one segment and there is no crossover case.
// Outputs are re-blocked into vectors of exactly block_size_out elements like advanceOutBuf (Filter.hs:516-523).
#include "records.cuh"

#include <deque>

namespace sdr {

// linear device buffer with a live region [rd, wr) in bytes
struct LinBuf {
    Ctx *c = nullptr;
    char *p = nullptr;
    size_t cap = 0, rd = 0, wr = 0;
    size_t size() const { return wr - rd; }
    int reserve(size_t more) {   // make room for `more` bytes after wr, keeping [rd, wr)
        if (wr + more <= cap) return SDR_OK;
        size_t live = size();
        if (live + more <= cap && rd >= live) {   // slide the live region to the front (regions do not overlap)
            if (live) SDR_CUDA(cudaMemcpyAsync(p, p + rd, live, cudaMemcpyDeviceToDevice, c->stream));
            rd = 0; wr = live;
            return SDR_OK;
        }
        size_t ncap = cap ? cap : ((size_t)1 << 20);
        while (ncap < 2 * (live + more)) ncap *= 2;
        char *np = nullptr;
        SDR_CUDA(cudaMalloc(&np, ncap));
        if (live) SDR_CUDA(cudaMemcpyAsync(np, p + rd, live, cudaMemcpyDeviceToDevice, c->stream));
        if (p) { SDR_CUDA(cudaStreamSynchronize(c->stream)); SDR_CUDA(cudaFree(p)); }
        p = np; cap = ncap; rd = 0; wr = live;
        return SDR_OK;
    }
    void consume(size_t bytes) { rd += bytes; if (rd == wr) rd = wr = 0; }
    void release() { if (p) cudaFree(p); p = nullptr; cap = rd = wr = 0; }
};

enum { P_FILTER, P_DECIM, P_RESAMP, P_FMDEMOD, P_CONVERT, P_SCALE };

}  // namespace sdr

using namespace sdr;

struct sdr_pipe {
    int kind = 0;
    Ctx *ctx = nullptr;
    FirRec *fir = nullptr;
    ResRec *res = nullptr;
    size_t in_eb = 4, out_eb = 4;
    long long block_out = 0;          // FIR kinds: elements per yielded vector
    LinBuf in;                        // FIR kinds: carried tail + pushed data
    LinBuf fifo;                      // produced, not yet popped / forwarded
    std::deque<long long> vec_lens;   // element-wise kinds: lengths of the queued vectors
    // resampler stream bookkeeping (global indices)
    long long k_next = 0;             // next output index
    long long pos = 0;                // global input index of in.rd
    long long n_total = 0;            // input elements pushed so far
    float scale_k = 1.0f;
    float *d_last = nullptr;          // fmDemod: previous sample (re, im), starts at 0 (Demod.hs:41)
    sdr_pipe *downstream = nullptr;
    const char *assert_name = "";
    // SDR_HOST_PINNED pushes: host-to-device copies of vectors that are contiguous on both sides are merged and
    // issued only when a launch needs the data (one DMA per output vector instead of one per input vector)
    long long batch_min = 0;          // launch only once this many new outputs are computable (0: one output vector)
    const char *pend_src = nullptr;
    char *pend_dst = nullptr;
    size_t pend_bytes = 0;
};

namespace sdr {

static bool is_fir_kind(int k) { return k == P_FILTER || k == P_DECIM || k == P_RESAMP; }

static int flush_pending(sdr_pipe *p) {
    if (p->pend_bytes) {
        SDR_CUDA(cudaMemcpyAsync(p->pend_dst, p->pend_src, p->pend_bytes, cudaMemcpyHostToDevice, p->ctx->stream));
        p->pend_bytes = 0; p->pend_src = nullptr; p->pend_dst = nullptr;
    }
    return SDR_OK;
}

static int pipe_push_dev(sdr_pipe *p, const void *d_src, long long n);

// hand everything that is complete to the connected stage (device to device, stream ordered)
static int forward(sdr_pipe *p) {
    if (!p->downstream) return SDR_OK;
    if (is_fir_kind(p->kind)) {
        long long have = (long long)(p->fifo.size() / p->out_eb);
        long long nb = have / p->block_out;
        if (nb > 0 && is_fir_kind(p->downstream->kind)) {
            // a FIR stage only sees the flat stream: hand it all complete vectors as one contiguous push
            SDR_TRY(pipe_push_dev(p->downstream, p->fifo.p + p->fifo.rd, nb * p->block_out));
            p->fifo.consume((size_t)(nb * p->block_out) * p->out_eb);
        } else {
            // element-wise stages yield one vector per awaited vector: keep the vector structure
            for (long long b = 0; b < nb; b++) {
                SDR_TRY(pipe_push_dev(p->downstream, p->fifo.p + p->fifo.rd, p->block_out));
                p->fifo.consume((size_t)p->block_out * p->out_eb);
            }
        }
    } else {
        while (!p->vec_lens.empty()) {
            long long n = p->vec_lens.front();
            p->vec_lens.pop_front();
            SDR_TRY(pipe_push_dev(p->downstream, p->fifo.p + p->fifo.rd, n));
            p->fifo.consume((size_t)n * p->out_eb);
        }
    }
    return SDR_OK;
}

// run whatever the stream now allows (FIR kinds); data already appended to p->in
static int process_fir(sdr_pipe *p, bool force = false) {
    long long have = (long long)(p->in.size() / p->in_eb);
    const long long fifo_have = (long long)(p->fifo.size() / p->out_eb);
    const long long batch = (force || p->batch_min < p->block_out) ? p->block_out : p->batch_min;
    if (p->kind == P_RESAMP) {
        ResRec &r = *p->res;
        long long total_out = (p->n_total * r.L >= r.T) ? (p->n_total * r.L - r.T) / r.M + 1 : 0;
        long long count = total_out - p->k_next;
        // lazy: launch only when the new outputs complete at least one output vector (fewer, larger launches;
        // invisible to the caller because vectors are only ever yielded whole)
        if (count > 0 && fifo_have + count >= p->block_out && (fifo_have + count >= batch)) {
            SDR_TRY(flush_pending(p));
            long long i_k = (p->k_next * r.M + r.L - 1) / r.L;   // ceil(k M / L): first sample of output k
            SDR_TRY(p->fifo.reserve((size_t)count * p->out_eb));
            Seg2 seg = {p->in.p + p->in.rd, have, nullptr, 0};
            SDR_TRY(r.run(seg, i_k - p->pos, (int)(p->k_next % r.ng), p->fifo.p + p->fifo.wr, count, false));
            p->fifo.wr += (size_t)count * p->out_eb;
            p->k_next = total_out;
            long long new_pos = (p->k_next * r.M + r.L - 1) / r.L;
            if (new_pos > p->n_total) new_pos = p->n_total;
            p->in.rd += (size_t)(new_pos - p->pos) * p->in_eb;   // the tail stays in place: no copy
            p->pos = new_pos;
        }
        return SDR_OK;
    }
    FirRec &f = *p->fir;
    long long count = (have >= f.T) ? (have - f.T) / f.D + 1 : 0;
    if (count > 0 && fifo_have + count >= p->block_out && (fifo_have + count >= batch)) {
        SDR_TRY(flush_pending(p));
        SDR_TRY(p->fifo.reserve((size_t)count * p->out_eb));
        Seg2 seg = {p->in.p + p->in.rd, have, nullptr, 0};
        SDR_TRY(f.run(seg, 0, p->fifo.p + p->fifo.wr, count, false));
        p->fifo.wr += (size_t)count * p->out_eb;
        p->in.rd += (size_t)(count * f.D) * p->in_eb;
    }
    return SDR_OK;
}

// copy n elements from host/device memory to dst on the ctx stream
static int fetch(sdr_pipe *p, void *d_dst, const void *src, size_t bytes, int mem) {
    if (!bytes) return SDR_OK;
    if (mem == SDR_HOST_PINNED) {
        if (p->pend_bytes && p->pend_src + p->pend_bytes == (const char *)src && p->pend_dst + p->pend_bytes == (char *)d_dst) {
            p->pend_bytes += bytes;   // extends the deferred copy
            return SDR_OK;
        }
        SDR_TRY(flush_pending(p));
        p->pend_src = (const char *)src; p->pend_dst = (char *)d_dst; p->pend_bytes = bytes;
        return SDR_OK;
    }
    SDR_TRY(flush_pending(p));
    if (mem == SDR_DEVICE) { SDR_CUDA(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyDeviceToDevice, p->ctx->stream)); return SDR_OK; }
    SDR_CUDA(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, p->ctx->stream));
    // pageable memory has been staged by the driver when the call returns; page-locked memory has not -- wait for
    // the copy so the caller's vector is free on return (FilterInternal.hs:68-71 pins for the call only)
    cudaPointerAttributes at;
    bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) SDR_CUDA(cudaStreamSynchronize(p->ctx->stream));
    return SDR_OK;
}

static int pipe_push_any(sdr_pipe *p, const void *src, long long n, int mem) {
    SDR_TRY(p->ctx->bind());
    if (is_fir_kind(p->kind)) {
        // the reference asserts every awaited vector holds at least numCoeffs samples (Filter.hs:544,586,691)
        long long need = (p->kind == P_RESAMP) ? (p->res->T + p->res->L - 1) / p->res->L : p->fir->T;
        if (n < need) return set_error(SDR_EPRECOND, "%s 1: input vector of %lld elements is shorter than numCoeffs (%lld)",
                                       p->assert_name, n, need);
        if (p->in.wr + (size_t)n * p->in_eb > p->in.cap) SDR_TRY(flush_pending(p));   // the buffer is about to move
        SDR_TRY(p->in.reserve((size_t)n * p->in_eb));
        SDR_TRY(fetch(p, p->in.p + p->in.wr, src, (size_t)n * p->in_eb, mem));
        p->in.wr += (size_t)n * p->in_eb;
        p->n_total += n;
        SDR_TRY(process_fir(p));
        return forward(p);
    }
    // element-wise kinds: one output vector per input vector
    long long n_out = n;
    const void *d_src = src;
    if (mem != SDR_DEVICE) {
        SDR_TRY(p->in.reserve((size_t)n * p->in_eb));
        SDR_TRY(fetch(p, p->in.p + p->in.wr, src, (size_t)n * p->in_eb, mem));
        SDR_TRY(flush_pending(p));
        d_src = p->in.p + p->in.wr;   // scratch use: the region is not kept
    }
    if (p->kind == P_CONVERT) {
        if (n & 1) return set_error(SDR_EINVAL, "convert pipe: odd byte count %lld (interleaved I/Q pairs expected)", n);
        n_out = n / 2;   // complex samples out
    }
    SDR_TRY(p->fifo.reserve((size_t)n_out * p->out_eb));
    void *d_dst = p->fifo.p + p->fifo.wr;
    if (p->kind == P_CONVERT) SDR_TRY(launch_convert_u8(p->ctx, (const uint8_t *)d_src, (float *)d_dst, n));
    else if (p->kind == P_SCALE) SDR_TRY(launch_scale(p->ctx, p->scale_k, (const float *)d_src, (float *)d_dst, n));
    else {
        SDR_TRY(launch_fm_demod_carry(p->ctx, p->d_last, (const float *)d_src, (float *)d_dst, n));
        if (n) SDR_CUDA(cudaMemcpyAsync(p->d_last, (const char *)d_src + (size_t)(n - 1) * 8, 8, cudaMemcpyDeviceToDevice, p->ctx->stream));
    }
    p->fifo.wr += (size_t)n_out * p->out_eb;
    p->vec_lens.push_back(n_out);
    return forward(p);
}

static int pipe_push_dev(sdr_pipe *p, const void *d_src, long long n) { return pipe_push_any(p, d_src, n, SDR_DEVICE); }

static int new_pipe(Ctx *c, int kind, sdr_pipe_t **out, sdr_pipe **p) {
    if (!c || !out) return set_error(SDR_EINVAL, "pipe constructor: bad argument");
    *out = nullptr;
    SDR_TRY(c->bind());
    sdr_pipe *h = new sdr_pipe();
    h->kind = kind; h->ctx = c; h->in.c = c; h->fifo.c = c;
    *p = h; *out = h;
    return SDR_OK;
}

}  // namespace sdr

extern "C" {

int sdr_pipe_fir_filter(sdr_filter_t *f, int block_size_out, sdr_pipe_t **out) {
    if (!f || block_size_out <= 0) return set_error(SDR_EINVAL, "sdr_pipe_fir_filter: bad argument");
    sdr_pipe *p;
    SDR_TRY(new_pipe(f->r.ctx, P_FILTER, out, &p));
    p->fir = &f->r; p->in_eb = p->out_eb = elem_bytes(f->r.cplx); p->block_out = block_size_out; p->assert_name = "filter";
    return SDR_OK;
}
int sdr_pipe_fir_decimator(sdr_decimator_t *d, int block_size_out, sdr_pipe_t **out) {
    if (!d || block_size_out <= 0) return set_error(SDR_EINVAL, "sdr_pipe_fir_decimator: bad argument");
    sdr_pipe *p;
    SDR_TRY(new_pipe(d->r.ctx, P_DECIM, out, &p));
    p->fir = &d->r; p->in_eb = p->out_eb = elem_bytes(d->r.cplx); p->block_out = block_size_out; p->assert_name = "decimate";
    return SDR_OK;
}
int sdr_pipe_fir_resampler(sdr_resampler_t *r, int block_size_out, sdr_pipe_t **out) {
    if (!r || block_size_out <= 0) return set_error(SDR_EINVAL, "sdr_pipe_fir_resampler: bad argument");
    sdr_pipe *p;
    SDR_TRY(new_pipe(r->r.ctx, P_RESAMP, out, &p));
    p->res = &r->r; p->in_eb = p->out_eb = elem_bytes(r->r.cplx); p->block_out = block_size_out; p->assert_name = "resample";
    return SDR_OK;
}
int sdr_pipe_fm_demod(sdr_ctx_t *ctx, sdr_pipe_t **out) {
    sdr_pipe *p;
    SDR_TRY(new_pipe(reinterpret_cast<Ctx *>(ctx), P_FMDEMOD, out, &p));
    p->in_eb = 8; p->out_eb = 4;
    SDR_CUDA(cudaMalloc(&p->d_last, 8));
    SDR_CUDA(cudaMemsetAsync(p->d_last, 0, 8, p->ctx->stream));
    return SDR_OK;
}
int sdr_pipe_convert_u8(sdr_ctx_t *ctx, sdr_pipe_t **out) {
    sdr_pipe *p;
    SDR_TRY(new_pipe(reinterpret_cast<Ctx *>(ctx), P_CONVERT, out, &p));
    p->in_eb = 1; p->out_eb = 8;
    return SDR_OK;
}
int sdr_pipe_scale(sdr_ctx_t *ctx, float factor, sdr_pipe_t **out) {
    sdr_pipe *p;
    SDR_TRY(new_pipe(reinterpret_cast<Ctx *>(ctx), P_SCALE, out, &p));
    p->in_eb = p->out_eb = 4; p->scale_k = factor;
    return SDR_OK;
}
int sdr_pipe_destroy(sdr_pipe_t *p) {
    if (!p) return SDR_OK;
    p->ctx->bind();
    cudaStreamSynchronize(p->ctx->stream);
    p->in.release(); p->fifo.release();
    if (p->d_last) cudaFree(p->d_last);
    delete p;
    return SDR_OK;
}

int sdr_pipe_push(sdr_pipe_t *p, const void *in, long long n, int mem) {
    if (!p || n < 0 || (n && !in) || mem < SDR_HOST || mem > SDR_HOST_PINNED)
        return set_error(SDR_EINVAL, "sdr_pipe_push: bad argument");
    if (n == 0) return SDR_OK;
    return pipe_push_any(p, in, n, mem);
}

int sdr_pipe_ready(sdr_pipe_t *p, int *n_blocks) {
    if (!p || !n_blocks) return set_error(SDR_EINVAL, "sdr_pipe_ready: bad argument");
    if (is_fir_kind(p->kind)) *n_blocks = (int)((long long)(p->fifo.size() / p->out_eb) / p->block_out);
    else *n_blocks = (int)p->vec_lens.size();
    return SDR_OK;
}

int sdr_pipe_next_len(sdr_pipe_t *p, long long *n) {
    if (!p || !n) return set_error(SDR_EINVAL, "sdr_pipe_next_len: bad argument");
    if (is_fir_kind(p->kind)) {
        if ((long long)(p->fifo.size() / p->out_eb) < p->block_out) return set_error(SDR_EAGAIN, "sdr_pipe_next_len: no complete output block yet");
        *n = p->block_out;
    } else {
        if (p->vec_lens.empty()) return set_error(SDR_EAGAIN, "sdr_pipe_next_len: no output vector yet");
        *n = p->vec_lens.front();
    }
    return SDR_OK;
}

int sdr_pipe_pop(sdr_pipe_t *p, void *out, long long *n_out, int mem) {
    if (!p || !out || mem < SDR_HOST || mem > SDR_HOST_PINNED) return set_error(SDR_EINVAL, "sdr_pipe_pop: bad argument");
    long long n;
    if (is_fir_kind(p->kind)) {
        n = p->block_out;
        if ((long long)(p->fifo.size() / p->out_eb) < n) return set_error(SDR_EAGAIN, "sdr_pipe_pop: no complete output block yet");
    } else {
        if (p->vec_lens.empty()) return set_error(SDR_EAGAIN, "sdr_pipe_pop: no output vector yet");
        n = p->vec_lens.front();
        p->vec_lens.pop_front();
    }
    SDR_TRY(p->ctx->bind());
    size_t bytes = (size_t)n * p->out_eb;
    if (bytes)
        SDR_CUDA(cudaMemcpyAsync(out, p->fifo.p + p->fifo.rd, bytes,
                                 mem == SDR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, p->ctx->stream));
    if (mem == SDR_HOST) SDR_CUDA(cudaStreamSynchronize(p->ctx->stream));
    p->fifo.consume(bytes);
    if (n_out) *n_out = n;
    return SDR_OK;
}

int sdr_pipe_sync(sdr_pipe_t *p) {
    if (!p) return set_error(SDR_EINVAL, "sdr_pipe_sync: null handle");
    SDR_TRY(p->ctx->bind());
    SDR_TRY(flush_pending(p));
    SDR_CUDA(cudaStreamSynchronize(p->ctx->stream));
    return SDR_OK;
}

int sdr_pipe_set_batch(sdr_pipe_t *p, long long min_outputs) {
    if (!p || min_outputs < 0) return set_error(SDR_EINVAL, "sdr_pipe_set_batch: bad argument");
    p->batch_min = min_outputs;
    if (is_fir_kind(p->kind) && min_outputs > 0) {
        // size both buffers for the batch once, instead of growing by doubling while the stream runs
        SDR_TRY(p->ctx->bind());
        long long in_per_out = (p->kind == P_RESAMP) ? (p->res->M + p->res->L - 1) / p->res->L : p->fir->D;
        long long taps = (p->kind == P_RESAMP) ? p->res->T : p->fir->T;
        SDR_TRY(p->in.reserve((size_t)(2 * (min_outputs + p->block_out) * in_per_out + taps) * p->in_eb));
        SDR_TRY(p->fifo.reserve((size_t)(2 * (min_outputs + p->block_out)) * p->out_eb));
    }
    return SDR_OK;
}

// pop every complete vector of `sink` into out[written...] with ONE copy (FIR kinds) / one copy per vector otherwise
static int drain(sdr_pipe *sink, void *out, long long out_capacity, int out_mem, long long *written) {
    if (is_fir_kind(sink->kind)) {
        long long nb = (long long)(sink->fifo.size() / sink->out_eb) / sink->block_out;
        if (nb == 0) return SDR_OK;
        long long n = nb * sink->block_out;
        if (*written + n > out_capacity) return set_error(SDR_EINVAL, "sdr_pipe_run: output capacity %lld too small", out_capacity);
        SDR_CUDA(cudaMemcpyAsync((char *)out + (size_t)*written * sink->out_eb, sink->fifo.p + sink->fifo.rd, (size_t)n * sink->out_eb,
                                 out_mem == SDR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, sink->ctx->stream));
        if (out_mem == SDR_HOST) SDR_CUDA(cudaStreamSynchronize(sink->ctx->stream));
        sink->fifo.consume((size_t)n * sink->out_eb);
        *written += n;
        return SDR_OK;
    }
    while (!sink->vec_lens.empty()) {
        if (*written + sink->vec_lens.front() > out_capacity)
            return set_error(SDR_EINVAL, "sdr_pipe_run: output capacity %lld too small", out_capacity);
        long long got = 0;
        SDR_TRY(sdr_pipe_pop(sink, (char *)out + (size_t)*written * sink->out_eb, &got, out_mem));
        *written += got;
    }
    return SDR_OK;
}

// runEffect $ each vectors >-> p >-> ... >-> sink >-> collect, as one native loop
int sdr_pipe_run(sdr_pipe_t *p, sdr_pipe_t *sink, const void *in, long long vec_len, long long n_vecs, int in_mem,
                 void *out, long long out_capacity, int out_mem, long long *n_out) {
    if (!p || !sink || vec_len <= 0 || n_vecs < 0 || (n_vecs && !in) || (out_capacity && !out) || !n_out ||
        in_mem < SDR_HOST || in_mem > SDR_HOST_PINNED || out_mem < SDR_HOST || out_mem > SDR_HOST_PINNED)
        return set_error(SDR_EINVAL, "sdr_pipe_run: bad argument");
    long long written = 0;
    SDR_TRY(p->ctx->bind());
    for (long long v = 0; v < n_vecs; v++) {
        SDR_TRY(pipe_push_any(p, (const char *)in + (size_t)(v * vec_len) * p->in_eb, vec_len, in_mem));
        SDR_TRY(drain(sink, out, out_capacity, out_mem, &written));
    }
    // end of input: run what the batching knob was still holding back, stage by stage
    for (sdr_pipe *q = p; q; q = q->downstream) {
        if (is_fir_kind(q->kind)) { SDR_TRY(process_fir(q, true)); SDR_TRY(forward(q)); }
        if (q == sink) break;
    }
    SDR_TRY(drain(sink, out, out_capacity, out_mem, &written));
    SDR_TRY(sdr_pipe_sync(p));
    *n_out = written;
    return SDR_OK;
}

int sdr_pipe_connect(sdr_pipe_t *src, sdr_pipe_t *dst) {
    if (!src || !dst || src == dst) return set_error(SDR_EINVAL, "sdr_pipe_connect: bad argument");
    if (src->ctx != dst->ctx) return set_error(SDR_EINVAL, "sdr_pipe_connect: stages live on different contexts");
    if (src->out_eb != dst->in_eb && !(dst->kind == P_CONVERT))
        return set_error(SDR_EINVAL, "sdr_pipe_connect: element types differ (%zu-byte out, %zu-byte in)", src->out_eb, dst->in_eb);
    if (is_fir_kind(src->kind) && is_fir_kind(dst->kind)) {
        long long need = (dst->kind == P_RESAMP) ? (dst->res->T + dst->res->L - 1) / dst->res->L : dst->fir->T;
        if (src->block_out < need)
            return set_error(SDR_EPRECOND, "%s 1: upstream vectors of %lld elements are shorter than numCoeffs (%lld)",
                             dst->assert_name, src->block_out, need);
    }
    src->downstream = dst;
    return SDR_OK;
}

}  // extern "C"
