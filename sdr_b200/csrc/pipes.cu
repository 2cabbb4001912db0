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

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>

#include <unistd.h>

namespace sdr {

// linear device buffer with a live region [rd, wr) in bytes
struct LinBuf {
    Ctx *c = nullptr;
    char *p = nullptr;
    size_t cap = 0, rd = 0, wr = 0;
    bool fixed = false;   // a persistent consumer writes into this buffer: it must neither move nor rewind
    bool borrowed = false; // the memory is the caller's output buffer for the duration of one sdr_pipe_run: not ours to move or free
    size_t size() const { return wr - rd; }
    int reserve(size_t more) {   // make room for `more` bytes after wr, keeping [rd, wr)
        if (wr + more <= cap) return SDR_OK;
        if (fixed) return set_error(SDR_ENOMEM, "buffer is held in place by a persistent consumer session (%zu more bytes wanted)", more);
        if (borrowed) return set_error(SDR_ENOMEM, "borrowed output buffer exhausted (%zu more bytes wanted)", more);
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
    void consume(size_t bytes) { rd += bytes; if (rd == wr && !fixed && !borrowed) rd = wr = 0; }
    // move the (small) live region so that it starts `lead` bytes past a 16-byte boundary at the front of the buffer;
    // skipped when source and destination would overlap
    int realign(size_t lead) {
        size_t live = size();
        if (rd == lead) return SDR_OK;
        if (lead + live > rd || lead + live > cap) return SDR_OK;
        if (live) SDR_CUDA(cudaMemcpyAsync(p + lead, p + rd, live, cudaMemcpyDeviceToDevice, c->stream));
        rd = lead; wr = lead + live;
        return SDR_OK;
    }
    void release() { if (p && !borrowed) cudaFree(p); p = nullptr; cap = rd = wr = 0; }
};

// Tracing aid.  SDR_B200_TRACE=1: synchronise around every stage step and print its wall time (serialises the stream).
// SDR_B200_TRACE=2: record CUDA events around every step without synchronising, print device-side durations and
// host-side issue times at the next sdr_pipe_sync.
static int trace_mode() { static const int m = getenv("SDR_B200_TRACE") ? atoi(getenv("SDR_B200_TRACE")) : 0; return m; }
struct TraceRec { const char *label; long long n; cudaEvent_t e0, e1; double host_us; };
static std::deque<TraceRec> &trace_log() { static std::deque<TraceRec> q; return q; }
struct TraceScope {
    const char *label; Ctx *c; long long n; std::chrono::steady_clock::time_point t0; cudaEvent_t e0 = nullptr;
    static bool active() { return trace_mode() == 1 || trace_mode() == 2; }   // 3 = run-level timing only
    TraceScope(const char *l, Ctx *ctx, long long count) : label(l), c(ctx), n(count) {
        if (!active() || !label) return;
        if (trace_mode() == 1) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->side); }
        else { cudaEventCreate(&e0); cudaEventRecord(e0, c->stream); }
        t0 = std::chrono::steady_clock::now();
    }
    ~TraceScope() {
        if (!active() || !label) return;
        if (trace_mode() == 1) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->side); }
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        if (trace_mode() == 1) { fprintf(stderr, "[sdr_b200 trace] %-18s n=%-10lld %9.1f us\n", label, n, us); return; }
        cudaEvent_t e1; cudaEventCreate(&e1); cudaEventRecord(e1, c->stream);
        trace_log().push_back({label, n, e0, e1, us});
    }
};
static void trace_dump() {
    if (trace_mode() != 2) return;
    cudaEvent_t first = trace_log().empty() ? nullptr : trace_log().front().e0;
    for (auto &r : trace_log()) {
        float ms = 0, at = 0;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventElapsedTime(&at, first, r.e0);
        fprintf(stderr, "[sdr_b200 trace] %-18s n=%-10lld device %9.1f us (starts at %9.1f us)  host issue %8.1f us\n", r.label, r.n,
                ms * 1e3, at * 1e3, r.host_us);
    }
    for (auto &r : trace_log()) { if (r.e0 != first) cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    if (first) cudaEventDestroy(first);
    trace_log().clear();
}

enum { P_FILTER, P_DECIM, P_RESAMP, P_FMDEMOD, P_CONVERT, P_SCALE, P_FMFRONT, P_DCBLOCK, P_U8DECIM, P_FMLOW };
enum { MEM_FWD = 100 };   // internal: a connected upstream stage hands over its FIFO region, read in place

}  // namespace sdr

using namespace sdr;

// A persistent-consumer session (kernels_fast.cu: k_dec_ring_persist): the resident kernel, its page-locked control block
// and where its input run and its outputs live.
struct PersistSession {
    bool open = false;
    void *ctl = nullptr; size_t ctl_bytes = 0;   // cudaHostAllocMapped: header + one done flag per run
    void *d_relay = nullptr;                     // device copy of (published, closed) kept current by the kernel's relay CTA
    cudaStream_t stream = nullptr;               // the consumer's own stream (the ctx stream stays free for everything else)
    cudaEvent_t ev = nullptr;
    const char *base = nullptr;                  // first byte of the session's in-place input run
    long long runs_total = 0, runs_popped = 0;   // capacity / runs whose outputs have been handed to the FIFO
    int run_samples = 0, halo_samples = 0, grid = 0;
    size_t fifo_base = 0;                        // FIFO offset of run 0's outputs
    long long capacity_bytes = 0;
    const char *kernel = "none";
    cudaEvent_t dbg0 = nullptr, dbg1 = nullptr;  // SDR_B200_PERSIST_DEBUG: device time of the resident kernel
    std::chrono::steady_clock::time_point t_open;
};

struct sdr_pipe {
    int kind = 0;
    Ctx *ctx = nullptr;
    FirRec *fir = nullptr;
    ResRec *res = nullptr;
    size_t in_eb = 4, out_eb = 4;
    long long block_out = 0;          // FIR kinds: elements per yielded vector
    LinBuf in;                        // FIR kinds: carried tail + pushed data
    LinBuf fifo;                      // produced, not yet popped / forwarded
    LinBuf fifo_own;                  // the stage's own FIFO while `fifo` is the caller's output buffer (sdr_pipe_run, device output)
    std::deque<long long> vec_lens;   // element-wise kinds: lengths of the queued vectors
    // resampler stream bookkeeping (global indices)
    long long k_next = 0;             // next output index
    long long pos = 0;                // global input index of in.rd
    long long n_total = 0;            // input elements pushed so far
    float scale_k = 1.0f;
    float *d_last = nullptr;          // fmDemod: previous sample (re, im), starts at 0 (Demod.hs:41)
    // fused FM front end: per-sub-tile boundary samples and scratch for the un-fused prefix / ragged end
    LinBuf bnd, scratch_x, scratch_y;
    int last_sel = 0;                 // which half of the double-buffered carried sample is current
    const char *last_kernel = "none";
    sdr_pipe *downstream = nullptr;
    sdr_pipe *upstream = nullptr;     // the stage connected in front of this one (unlinked when either side is destroyed)
    bool ext_fwd = false;             // the in-place run lies in the upstream stage's FIFO (held there, see fifo_reserve)
    long long skip = 0;               // input elements still to drop: a decimation factor larger than the tap count skips
                                      // past the resident data (the reference's VG.drop (count*D - len), Filter.hs:610)
    const char *assert_name = "";
    // SDR_HOST_PINNED pushes: host-to-device copies of vectors that are contiguous on both sides are merged and
    // issued only when a launch needs the data (one DMA per output vector instead of one per input vector)
    // full-duplex drain (sdr_pipe_run, pinned output): device-to-host copies ride the side stream so they overlap the
    // next batch's host-to-device copies; the FIFO may not be rewritten before the copy has finished
    cudaEvent_t ev_out_ready = nullptr, ev_out_done = nullptr;
    bool d2h_outstanding = false;
    long long batch_min = 0;          // launch only once this many new outputs are computable (0: one output vector)
    const char *pend_src = nullptr;
    char *pend_dst = nullptr;
    size_t pend_bytes = 0;
    // zero-copy input: a run of device vectors (SDR_DEVICE_HELD pushes adjacent in memory, or the FIFO region a connected
    // upstream stage hands over) that logically FOLLOWS the live region of `in` and is read in place by the next launch
    const char *ext_p = nullptr;
    size_t ext_bytes = 0;
    long long block_r = 0;            // fused low-rate stage: the resampler stage's own vector length (its yields gate the filter)
    long long persist_max = 0;        // > 0: in-place runs are consumed by a persistent kernel, sessions of at most this many samples
    int persist_idle_strikes = 0;     // consecutive sessions that wound down idle without finishing a single run
    PersistSession ps;
};

namespace sdr {

static bool is_fir_kind(int k) { return k == P_FILTER || k == P_DECIM || k == P_RESAMP || k == P_FMFRONT || k == P_U8DECIM || k == P_FMLOW; }
static bool is_resamp_kind(int k) { return k == P_RESAMP || k == P_FMLOW; }
static bool is_byte_fed(int k) { return k == P_FMFRONT || k == P_U8DECIM; }

static int fifo_writable(sdr_pipe *p) {
    if (p->d2h_outstanding) { SDR_CUDA(cudaStreamWaitEvent(p->ctx->stream, p->ev_out_done, 0)); p->d2h_outstanding = false; }
    return SDR_OK;
}

static int flush_pending(sdr_pipe *p) {
    if (p->pend_bytes) {
        TraceScope tr("h2d", p->ctx, (long long)p->pend_bytes);
        SDR_CUDA(cudaMemcpyAsync(p->pend_dst, p->pend_src, p->pend_bytes, cudaMemcpyHostToDevice, p->ctx->stream));
        p->pend_bytes = 0; p->pend_src = nullptr; p->pend_dst = nullptr;
    }
    return SDR_OK;
}

static long long in_elems(const sdr_pipe *p) { return (long long)((p->in.size() + p->ext_bytes) / p->in_eb); }
// the resident input stream as (carried tail, in-place run); one segment when there is no tail
static Seg2 input_seg(const sdr_pipe *p) {
    Seg2 s = {p->in.p + p->in.rd, (long long)(p->in.size() / p->in_eb), p->ext_p, (long long)(p->ext_bytes / p->in_eb)};
    if (s.na == 0 && s.nb > 0) { s.a = s.b; s.na = s.nb; s.b = nullptr; s.nb = 0; }
    return s;
}
static Seg2 seg_advance(Seg2 s, long long first, size_t eb) {
    if (first >= s.na) { s.a = (const char *)s.b + (size_t)(first - s.na) * eb; s.na = s.nb - (first - s.na); s.b = nullptr; s.nb = 0;
                         if (s.na < 0) s.na = 0; }
    else               { s.a = (const char *)s.a + (size_t)first * eb; s.na -= first; }
    return s;
}
// copy what is left of the in-place run behind the carried tail: afterwards the stage references only its own memory
static int materialize_ext(sdr_pipe *p) {
    if (!p->ext_bytes) return SDR_OK;
    SDR_TRY(flush_pending(p));
    SDR_TRY(p->in.reserve(p->ext_bytes));
    SDR_CUDA(cudaMemcpyAsync(p->in.p + p->in.wr, p->ext_p, p->ext_bytes, cudaMemcpyDeviceToDevice, p->ctx->stream));
    p->in.wr += p->ext_bytes;
    p->ext_p = nullptr; p->ext_bytes = 0; p->ext_fwd = false;
    return SDR_OK;
}
// A connected stage whose launch threshold (sdr_pipe_set_batch) has not been reached keeps the vectors its upstream stage
// handed over WHERE THEY ARE, in that stage's FIFO (round 1 copied every hand-over into the stage: 8 MB per 32 MiB push of
// the FM chain, a quarter of the front-end kernel's time).  The upstream stage then neither rewinds nor moves its FIFO: it
// appends behind the region, so the next hand-over is adjacent and extends the run, until it runs out of room -- only then
// does the downstream stage take what it still references into its own buffer.
static bool fifo_leased(const sdr_pipe *p) { return p->downstream && p->downstream->ext_fwd && p->downstream->ext_bytes > 0; }
// sdr_pipe_run with a device output buffer lets the sink stage produce STRAIGHT INTO that buffer (no copy of the yielded
// vectors): the stage's FIFO is the caller's memory for the duration of the call.  fifo_return gives it back: what has not
// been yielded yet (the partial output block, or everything still undrained when the buffer runs out) moves into the
// stage's own FIFO.
static int fifo_return(sdr_pipe *p) {
    if (!p->fifo.borrowed) return SDR_OK;
    LinBuf b = p->fifo;
    p->fifo = p->fifo_own;
    p->fifo_own = LinBuf();
    p->fifo.rd = p->fifo.wr = 0;
    const size_t live = b.size();
    if (live) {
        SDR_TRY(p->fifo.reserve(live));
        SDR_CUDA(cudaMemcpyAsync(p->fifo.p, b.p + b.rd, live, cudaMemcpyDeviceToDevice, p->ctx->stream));
        p->fifo.wr = live;
    }
    return SDR_OK;
}
static int fifo_reserve(sdr_pipe *p, size_t bytes) {
    if (p->fifo.borrowed && p->fifo.wr + bytes > p->fifo.cap) SDR_TRY(fifo_return(p));   // the caller's buffer is too small to work in
    if (fifo_leased(p) && p->fifo.wr + bytes > p->fifo.cap) SDR_TRY(materialize_ext(p->downstream));
    if (!fifo_leased(p) && p->fifo.rd == p->fifo.wr && !p->fifo.fixed && !p->fifo.borrowed) p->fifo.rd = p->fifo.wr = 0;
    return p->fifo.reserve(bytes);
}
// `bytes` at the front of the FIFO have been handed to the downstream stage
static void fifo_forwarded(sdr_pipe *p, size_t bytes) {
    if (fifo_leased(p)) p->fifo.rd += bytes;   // no rewind: the region is still being referenced
    else p->fifo.consume(bytes);
}
// After a launch of a fused byte-fed stage: the few samples it left over.  When they lie in the in-place run and the next
// launch can read them there (16-byte aligned start, nothing carried in the stage's own buffer) they STAY in the caller's
// vector -- his until sdr_pipe_sync -- and the next adjacent push simply extends the run: no copy between two launches, so
// consecutive kernels are neighbours in the stream and overlap their launch latency (programmatic dependent launch).
// Otherwise (and always for a connected upstream stage's FIFO region, see sdr_pipe_push) the tail moves into the stage.
static int park_tail(sdr_pipe *p) {
    if (p->ext_bytes && !p->ext_fwd && p->in.size() == 0 && (((uintptr_t)p->ext_p) & 15) == 0) return SDR_OK;
    return materialize_ext(p);
}
// drop `bytes` from the front of the resident stream (tail first, then the in-place run)
static void consume_input(sdr_pipe *p, size_t bytes) {
    const size_t live = p->in.size();
    if (bytes < live) { p->in.rd += bytes; return; }
    p->in.rd = p->in.wr = 0;
    const size_t rest = bytes - live;
    p->ext_p += rest; p->ext_bytes -= rest;
    if (p->ext_bytes == 0) p->ext_p = nullptr;
}

// ---- persistent consumer sessions ---------------------------------------------------------------------------------------
// layout of kernels_fast.cu's PersistCtl: one cache line per writer (host: bytes 0..63, device: 64.., flags from 128)
struct PersistHdr { volatile long long published_bytes; volatile int closed; int pad0_[13]; volatile int error; int pad1_[15]; };
static_assert(sizeof(PersistHdr) == 128, "control block layout");
static volatile unsigned int *persist_flags(PersistSession &S) { return (volatile unsigned int *)((char *)S.ctl + sizeof(PersistHdr)); }

// hand the outputs of every run that has completed (in order) to the FIFO
static void persist_poll(sdr_pipe *p) {
    PersistSession &S = p->ps;
    if (!S.open) return;
    volatile unsigned int *done = persist_flags(S);
    long long r = S.runs_popped;
    while (r < S.runs_total && done[r]) r++;
    if (r != S.runs_popped) {
        S.runs_popped = r;
        p->fifo.wr = S.fifo_base + (size_t)r * (size_t)(S.run_samples / p->fir->D) * p->out_eb;
    }
}

// end the session: the kernel finishes every completely published run and exits; what is left of the in-place run (the
// carried tail and an incomplete run) goes back to the ordinary launch path
static int persist_close(sdr_pipe *p) {
    PersistSession &S = p->ps;
    if (!S.open) return SDR_OK;
    PersistHdr *h = (PersistHdr *)S.ctl;
    static const bool dbg = getenv("SDR_B200_PERSIST_DEBUG") != nullptr;
    if (dbg) { persist_poll(p); fprintf(stderr, "[persist] closing: published %lld bytes, %lld of %lld runs popped so far\n", (long long)h->published_bytes, S.runs_popped, S.runs_total); }
    __sync_synchronize();
    h->closed = 1;
    __sync_synchronize();
    auto t_close = std::chrono::steady_clock::now();
    SDR_CUDA(cudaStreamSynchronize(S.stream));
    if (dbg && S.dbg0) {
        float ms = 0;
        cudaEventElapsedTime(&ms, S.dbg0, S.dbg1);
        fprintf(stderr, "[persist] resident kernel %.1f us on the device; host: open -> close request %.1f us, close request -> kernel gone %.1f us\n", ms * 1e3,
                std::chrono::duration<double, std::micro>(t_close - S.t_open).count(),
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_close).count());
    }
    persist_poll(p);
    S.open = false;
    p->fifo.fixed = false;
    const long long pub_samples = h->published_bytes / (long long)p->in_eb;
    const long long runs_final = pub_samples >= S.halo_samples ? (pub_samples - S.halo_samples) / S.run_samples : 0;
    // h->error: a warp waited more than 3 s for input and the kernel wound itself down (an idle stream must not keep a
    // spinning kernel resident for ever).  Not a failure: the runs it did finish count, the rest is computed again.
    if (!h->error && S.runs_popped != (runs_final < S.runs_total ? runs_final : S.runs_total))
        return set_error(SDR_ECUDA, "persistent consumer: %lld runs completed, %lld expected", S.runs_popped, runs_final);
    // Two sessions in a row that timed out without finishing a run: the host does not get to publish while the kernel is
    // resident (e.g. a profiler that serialises launches) -- the mode cannot work there, the launch path takes over.
    if (h->error && S.runs_popped == 0) { if (++p->persist_idle_strikes >= 2) p->persist_max = 0; }
    else p->persist_idle_strikes = 0;
    consume_input(p, (size_t)S.runs_popped * (size_t)S.run_samples * p->in_eb);
    return SDR_OK;
}

// open a session on the in-place run if there is none, publish what has arrived; false: the ordinary path must take over
static int persist_try(sdr_pipe *p, bool *taken) {
    *taken = false;
    PersistSession &S = p->ps;
    FirRec &f = *p->fir;
    if (!S.open) {
        if (!p->ext_bytes || p->skip != 0) return SDR_OK;
        SDR_TRY(flush_pending(p));
        SDR_TRY(fifo_writable(p));
        if (p->in.size() != 0) {
            // bridge: the windows that START in the carried tail straddle the stage's own buffer and the in-place run; one
            // small ordinary launch computes them, after which the stream continues inside the run alone
            const long long na = (long long)(p->in.size() / p->in_eb), nb = (long long)(p->ext_bytes / p->in_eb);
            const long long m1 = (na + f.D - 1) / f.D;
            if (m1 * f.D - na + f.T > nb) return SDR_OK;                      // not enough of the run yet
            SDR_TRY(fifo_reserve(p, (size_t)m1 * p->out_eb));
            SDR_TRY(launch_fir_generic(p->ctx, f.cplx, f.T, f.D, f.d_taps, input_seg(p), p->fifo.p + p->fifo.wr, m1));
            p->fifo.wr += (size_t)m1 * p->out_eb;
            consume_input(p, (size_t)(m1 * f.D) * p->in_eb);
        }
        if ((((uintptr_t)p->ext_p) & 15) != 0) return SDR_OK;
        if (!S.stream) {
            SDR_CUDA(cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking));
            SDR_CUDA(cudaEventCreateWithFlags(&S.ev, cudaEventDisableTiming));
        }
        // geometry of the kernel that will serve the shape
        if (!dec_persist_geometry(f.T, f.D, f.cplx, &S.run_samples, &S.halo_samples)) return SDR_OK;
        const long long max_samples = p->persist_max;
        const long long out_bytes = (max_samples / f.D + 1) * (long long)p->out_eb;
        SDR_TRY(fifo_reserve(p, (size_t)out_bytes));
        const long long runs_total = max_samples / S.run_samples;
        const size_t need = sizeof(PersistHdr) + (size_t)(runs_total + 1) * 4;
        if (need > S.ctl_bytes) {
            if (S.ctl) SDR_CUDA(cudaFreeHost(S.ctl));
            S.ctl = nullptr; S.ctl_bytes = 0;
            SDR_CUDA(cudaHostAlloc(&S.ctl, need, cudaHostAllocMapped));
            S.ctl_bytes = need;
        }
        memset(S.ctl, 0, need);
        S.runs_total = runs_total; S.runs_popped = 0;
        if (!S.d_relay) SDR_CUDA(cudaMalloc(&S.d_relay, 64));
        S.base = p->ext_p; S.fifo_base = p->fifo.wr;
        S.capacity_bytes = runs_total * (long long)S.run_samples * (long long)p->in_eb + (long long)S.halo_samples * (long long)p->in_eb;
        __sync_synchronize();
        void *d_ctl = nullptr;
        SDR_CUDA(cudaHostGetDevicePointer(&d_ctl, S.ctl, 0));
        // everything enqueued on the ctx stream so far (what produced the FIFO's contents, tail copies) precedes the consumer
        SDR_CUDA(cudaEventRecord(S.ev, p->ctx->stream));
        SDR_CUDA(cudaStreamWaitEvent(S.stream, S.ev, 0));
        static const bool dbg_t = getenv("SDR_B200_PERSIST_DEBUG") != nullptr;
        if (dbg_t) {
            if (!S.dbg0) { SDR_CUDA(cudaEventCreate(&S.dbg0)); SDR_CUDA(cudaEventCreate(&S.dbg1)); }
            SDR_CUDA(cudaEventRecord(S.dbg0, S.stream));
            S.t_open = std::chrono::steady_clock::now();
        }
        SDR_TRY(launch_dec_persist(p->ctx, f.T, f.D, f.cplx, f.d_taps, S.base, p->fifo.p + S.fifo_base, d_ctl, S.d_relay, runs_total, S.stream,
                                   &S.grid, &S.kernel));
        if (dbg_t) SDR_CUDA(cudaEventRecord(S.dbg1, S.stream));
        p->fifo.fixed = true;
        S.open = true;
        f.last_kernel = S.kernel;
    }
    if (p->ext_p != S.base || (long long)p->ext_bytes > S.capacity_bytes) { SDR_TRY(persist_close(p)); return SDR_OK; }
    PersistHdr *h = (PersistHdr *)S.ctl;
    __sync_synchronize();                    // the vectors were complete before they were pushed (SDR_DEVICE_HELD contract)
    h->published_bytes = (long long)p->ext_bytes;
    *taken = true;
    return SDR_OK;
}

static int pipe_push_dev(sdr_pipe *p, const void *d_src, long long n);
static int pipe_push_any(sdr_pipe *p, const void *src, long long n, int mem, long long n_vecs = 1);

// hand everything that is complete to the connected stage (device to device, stream ordered)
static int forward(sdr_pipe *p) {
    if (!p->downstream) return SDR_OK;
    if (is_fir_kind(p->kind)) {
        long long have = (long long)(p->fifo.size() / p->out_eb);
        long long nb = have / p->block_out;
        if (nb > 0 && is_fir_kind(p->downstream->kind)) {
            // a FIR stage only sees the flat stream: hand it all complete vectors as one contiguous push
            SDR_TRY(pipe_push_any(p->downstream, p->fifo.p + p->fifo.rd, nb * p->block_out, MEM_FWD));
            fifo_forwarded(p, (size_t)(nb * p->block_out) * p->out_eb);
        } else if (nb > 0) {
            // element-wise stages yield one vector per awaited vector: one launch over all of them, but the vector
            // structure (nb vectors of block_out elements) is kept in the stage's queue
            SDR_TRY(pipe_push_any(p->downstream, p->fifo.p + p->fifo.rd, nb * p->block_out, SDR_DEVICE, nb));
            p->fifo.consume((size_t)(nb * p->block_out) * p->out_eb);
        }
    } else {
        if (is_fir_kind(p->downstream->kind) && p->vec_lens.size() > 1) {
            // a FIR stage only sees the flat stream: hand it all queued vectors as one contiguous push (each must
            // still satisfy the stage's minimum-length precondition, checked on the shortest)
            long long total = 0, shortest = p->vec_lens.front();
            for (long long n : p->vec_lens) { total += n; if (n < shortest) shortest = n; }
            sdr_pipe *d = p->downstream;
            long long need = is_resamp_kind(d->kind) ? (d->res->T + d->res->L - 1) / d->res->L : d->fir->T;
            if (shortest >= need) {
                p->vec_lens.clear();
                SDR_TRY(pipe_push_any(d, p->fifo.p + p->fifo.rd, total, MEM_FWD));
                fifo_forwarded(p, (size_t)total * p->out_eb);
                return SDR_OK;
            }
        }
        while (!p->vec_lens.empty()) {
            long long n = p->vec_lens.front();
            p->vec_lens.pop_front();
            SDR_TRY(pipe_push_any(p->downstream, p->fifo.p + p->fifo.rd, n, is_fir_kind(p->downstream->kind) ? (int)MEM_FWD : (int)SDR_DEVICE));
            fifo_forwarded(p, (size_t)n * p->out_eb);
        }
    }
    return SDR_OK;
}

// un-fused convert -> decimate -> demod of outputs [first, first + n) of the resident byte stream; `slot` selects one
// of two scratch regions so the complex outputs of an earlier call stay readable.  *last = its final complex output.
static int fm_unfused(sdr_pipe *p, long long first, long long n, const float *d_carry, float *d_out, int slot,
                      const float **last) {
    FirRec &f = *p->fir;
    const long long n_s = (n - 1) * f.D + f.T;
    // region 0 (the <= 3-output alignment prefix) is a fixed 4 KB, region 1 (the ragged end) follows it
    const size_t xo = slot ? 4096 : 0, yo = slot ? 4096 : 0;
    p->scratch_x.rd = p->scratch_x.wr = 0; p->scratch_y.rd = p->scratch_y.wr = 0;
    SDR_TRY(p->scratch_x.reserve(4096 + (size_t)n_s * 8 + 256));
    SDR_TRY(p->scratch_y.reserve(4096 + (size_t)n * 8 + 256));
    float *x = (float *)(p->scratch_x.p + xo), *y = (float *)(p->scratch_y.p + yo);
    SDR_TRY(launch_convert_u8(p->ctx, (const uint8_t *)(p->in.p + p->in.rd) + 2 * first * f.D, x, 2 * n_s));
    Seg2 seg = {x, n_s, nullptr, 0};
    SDR_TRY(f.run(seg, 0, y, n, false));
    SDR_TRY(launch_fm_demod_carry(p->ctx, d_carry, y, d_out, n));
    *last = y + 2 * (n - 1);
    return SDR_OK;
}

// P.map convert >-> firDecimator >-> fmDemod as one stage.  The fused kernel covers every output (ragged last
// sub-tile and unaligned FIFO cursor included); shapes without a tuned kernel run the three stages un-fused.
static int process_fm_front(sdr_pipe *p, long long fifo_have, long long batch) {
    FirRec &f = *p->fir;
    const long long have = in_elems(p) / 2;   // IQ pairs resident (carried tail + in-place run)
    long long count = (have >= f.T) ? (have - f.T) / f.D + 1 : 0;
    if (!(count > 0 && fifo_have + count >= p->block_out && fifo_have + count >= batch)) return SDR_OK;
    SDR_TRY(flush_pending(p));
    TraceScope tr("fm_front", p->ctx, count);
    SDR_TRY(fifo_writable(p));
    SDR_TRY(fifo_reserve(p, (size_t)count * 4));
    float *out = (float *)(p->fifo.p + p->fifo.wr);
    p->bnd.rd = p->bnd.wr = 0;
    SDR_TRY(p->bnd.reserve((size_t)(count / 256 + 2) * 16));
    // the carried sample is double buffered: the kernel's boundary pass reads the old one, its main loop writes the new one;
    // the word behind the two samples is the kernel's ticket counter
    float *carry_in = p->d_last + 2 * p->last_sel, *carry_out = p->d_last + 2 * (p->last_sel ^ 1);
    long long done = 0;
    const char *name = nullptr;
    Seg2 seg = input_seg(p);   // element = byte
    SDR_TRY(launch_fm_front(p->ctx, f.T, f.D, f.d_taps, f.symmetric, (const uint8_t *)seg.a, seg.nb ? seg.na / 2 : have, (const uint8_t *)seg.b,
                            have, out, count, (float2 *)p->bnd.p, (long long)(p->bnd.cap / 16), (const float2 *)carry_in, (float2 *)carry_out,
                            (unsigned int *)(p->d_last + 4), &done, &name));
    p->last_kernel = name;
    if (done < count) {
        SDR_TRY(materialize_ext(p));   // the un-fused stages read one contiguous segment
        const float *last = nullptr;
        SDR_TRY(fm_unfused(p, 0, count, carry_in, out, 1, &last));
        SDR_CUDA(cudaMemcpyAsync(carry_out, last, 8, cudaMemcpyDeviceToDevice, p->ctx->stream));
    }
    p->last_sel ^= 1;
    p->fifo.wr += (size_t)count * 4;
    {
        const long long adv = count * f.D < have ? count * f.D : have;
        p->skip += 2 * (count * f.D - adv);
        consume_input(p, (size_t)adv * 2);
    }
    SDR_TRY(park_tail(p));
    if (p->in.size() <= (1u << 16) && p->in.rd > p->in.cap / 4) SDR_TRY(p->in.realign(0));
    return SDR_OK;
}

// P.map convert >-> firDecimator as one stage (complex outputs): the fused kernel without the discriminator; shapes
// without a tuned kernel convert into scratch and run the decimator record on it.
static int process_u8_decim(sdr_pipe *p, long long fifo_have, long long batch) {
    FirRec &f = *p->fir;
    const long long have = in_elems(p) / 2;   // IQ pairs resident
    long long count = (have >= f.T) ? (have - f.T) / f.D + 1 : 0;
    if (!(count > 0 && fifo_have + count >= p->block_out && fifo_have + count >= batch)) return SDR_OK;
    SDR_TRY(flush_pending(p));
    TraceScope tr("u8_decimator", p->ctx, count);
    SDR_TRY(fifo_writable(p));
    SDR_TRY(fifo_reserve(p, (size_t)count * 8));
    float *out = (float *)(p->fifo.p + p->fifo.wr);
    long long done = 0;
    const char *name = nullptr;
    Seg2 seg = input_seg(p);
    SDR_TRY(launch_dec_u8(p->ctx, f.T, f.D, f.d_taps, f.symmetric, (const uint8_t *)seg.a, seg.nb ? seg.na / 2 : have, (const uint8_t *)seg.b, have,
                          out, count, &done, &name));
    p->last_kernel = name;
    if (done < count) {
        SDR_TRY(materialize_ext(p));
        const long long n_s = (count - 1) * f.D + f.T;
        p->scratch_x.rd = p->scratch_x.wr = 0;
        SDR_TRY(p->scratch_x.reserve((size_t)n_s * 8 + 256));
        SDR_TRY(launch_convert_u8(p->ctx, (const uint8_t *)(p->in.p + p->in.rd), (float *)p->scratch_x.p, 2 * n_s));
        Seg2 one = {p->scratch_x.p, n_s, nullptr, 0};
        SDR_TRY(f.run(one, 0, out, count, false));
    }
    p->fifo.wr += (size_t)count * 8;
    {
        const long long adv = count * f.D < have ? count * f.D : have;
        p->skip += 2 * (count * f.D - adv);
        consume_input(p, (size_t)adv * 2);
    }
    SDR_TRY(park_tail(p));
    if (p->in.size() <= (1u << 16) && p->in.rd > p->in.cap / 4) SDR_TRY(p->in.realign(0));
    return SDR_OK;
}

// firResampler >-> firFilter >-> P.map (* k) as one stage (fm.hs:38-40).  The resampler stage of the un-fused chain only
// yields whole vectors of block_r elements, so the filter sees R_avail = floor(R_total / block_r) * block_r resampled
// samples and can produce R_avail - numCoeffsF + 1 outputs: the fused stage yields exactly those.
static int process_fm_low(sdr_pipe *p, long long fifo_have, long long batch) {
    ResRec &r = *p->res;
    FirRec &f = *p->fir;
    const long long r_total = (p->n_total * r.L >= r.T) ? (p->n_total * r.L - r.T) / r.M + 1 : 0;
    const long long r_avail = (r_total / p->block_r) * p->block_r;
    const long long z_total = r_avail >= f.T ? r_avail - f.T + 1 : 0;
    const long long count = z_total - p->k_next;
    if (!(count > 0 && fifo_have + count >= p->block_out && fifo_have + count >= batch)) return SDR_OK;
    SDR_TRY(flush_pending(p));
    TraceScope tr("fm_lowrate", p->ctx, count);
    SDR_TRY(fifo_writable(p));
    SDR_TRY(fifo_reserve(p, (size_t)count * 4));
    float *out = (float *)(p->fifo.p + p->fifo.wr);
    const long long i_k = (p->k_next * r.M + r.L - 1) / r.L;   // first sample of resampler output k_next
    Seg2 seg = seg_advance(input_seg(p), i_k - p->pos, 4);
    long long done = 0;
    const char *name = nullptr;
    SDR_TRY(launch_fm_lowrate(p->ctx, r.L, r.M, r.n_taps, r.h_plain.data(), f.T, f.h_taps.data(), p->scale_k, seg, p->k_next, out, count, &done, &name));
    p->last_kernel = name;
    if (done < count) {   // no fused kernel for the shape: the three stages one after the other through scratch
        const long long nr = count + f.T - 1;
        p->scratch_x.rd = p->scratch_x.wr = 0; p->scratch_y.rd = p->scratch_y.wr = 0;
        SDR_TRY(p->scratch_x.reserve((size_t)nr * 4 + 256));
        SDR_TRY(p->scratch_y.reserve((size_t)count * 4 + 256));
        SDR_TRY(r.run(seg, 0, (int)(p->k_next % r.ng), p->scratch_x.p, nr, false));
        Seg2 one = {p->scratch_x.p, nr, nullptr, 0};
        SDR_TRY(f.run(one, 0, p->scratch_y.p, count, false));
        SDR_TRY(launch_scale(p->ctx, p->scale_k, (const float *)p->scratch_y.p, out, count));
    }
    p->fifo.wr += (size_t)count * 4;
    p->k_next += count;
    long long new_pos = (p->k_next * r.M + r.L - 1) / r.L;
    if (new_pos > p->n_total) new_pos = p->n_total;
    consume_input(p, (size_t)(new_pos - p->pos) * 4);
    p->pos = new_pos;
    SDR_TRY(materialize_ext(p));
    if (p->in.size() <= (1u << 16) && p->in.rd > p->in.cap / 4) SDR_TRY(p->in.realign(0));
    return SDR_OK;
}

// run whatever the stream now allows (FIR kinds); data already appended to p->in
static int process_fir(sdr_pipe *p, bool force = false) {
    long long have = in_elems(p);
    long long fifo_have = (long long)(p->fifo.size() / p->out_eb);
    const long long batch = (force || p->batch_min < p->block_out) ? p->block_out : p->batch_min;
    if (p->kind == P_RESAMP) {
        ResRec &r = *p->res;
        long long total_out = (p->n_total * r.L >= r.T) ? (p->n_total * r.L - r.T) / r.M + 1 : 0;
        long long count = total_out - p->k_next;
        // lazy: launch only when the new outputs complete at least one output vector (fewer, larger launches;
        // invisible to the caller because vectors are only ever yielded whole)
        if (count > 0 && fifo_have + count >= p->block_out && (fifo_have + count >= batch)) {
            SDR_TRY(flush_pending(p));
            TraceScope tr("resampler", p->ctx, count);
            long long i_k = (p->k_next * r.M + r.L - 1) / r.L;   // ceil(k M / L): first sample of output k
            SDR_TRY(fifo_writable(p));
            SDR_TRY(fifo_reserve(p, (size_t)count * p->out_eb));
            SDR_TRY(r.run(input_seg(p), i_k - p->pos, (int)(p->k_next % r.ng), p->fifo.p + p->fifo.wr, count, false));
            p->fifo.wr += (size_t)count * p->out_eb;
            p->k_next = total_out;
            long long new_pos = (p->k_next * r.M + r.L - 1) / r.L;
            if (new_pos > p->n_total) new_pos = p->n_total;
            consume_input(p, (size_t)(new_pos - p->pos) * p->in_eb);
            p->pos = new_pos;
            SDR_TRY(materialize_ext(p));
            if (!r.cplx && p->in.size() <= (1u << 16)) {
                // the tuned kernel starts at the next cycle boundary (output index multiple of ng) with a 16-byte
                // aligned window: park the short tail so that this start lands on a 16-byte boundary
                long long kb = ((p->k_next + r.ng - 1) / r.ng) * r.ng;
                long long s0 = (kb * r.M + r.L - 1) / r.L - p->pos;   // floats from the tail start to that window
                SDR_TRY(p->in.realign((size_t)((4 - (s0 & 3)) & 3) * 4));
            }
        }
        return SDR_OK;
    }
    if (p->kind == P_FMFRONT) return process_fm_front(p, fifo_have, batch);
    if (p->kind == P_U8DECIM) return process_u8_decim(p, fifo_have, batch);
    if (p->kind == P_FMLOW) return process_fm_low(p, fifo_have, batch);
    FirRec &f = *p->fir;
    if (p->persist_max > 0) {
        if (!force) {
            bool taken = false;
            SDR_TRY(persist_try(p, &taken));
            if (taken) return SDR_OK;            // the resident kernel consumes the in-place run as it is published
        }
        SDR_TRY(persist_close(p));               // flush, or the run no longer qualifies: finish with ordinary launches
        have = in_elems(p);
        fifo_have = (long long)(p->fifo.size() / p->out_eb);
    }
    long long count = (have >= f.T) ? (have - f.T) / f.D + 1 : 0;
    if (count > 0 && fifo_have + count >= p->block_out && (fifo_have + count >= batch)) {
        SDR_TRY(flush_pending(p));
        TraceScope tr(p->kind == P_FILTER ? "filter" : "decimator", p->ctx, count);
        SDR_TRY(fifo_writable(p));
        SDR_TRY(fifo_reserve(p, (size_t)count * p->out_eb));
        SDR_TRY(f.run(input_seg(p), 0, p->fifo.p + p->fifo.wr, count, false));
        p->fifo.wr += (size_t)count * p->out_eb;
        {
            const long long adv = count * f.D < have ? count * f.D : have;
            p->skip += count * f.D - adv;
            consume_input(p, (size_t)adv * p->in_eb);
        }
        SDR_TRY(materialize_ext(p));
        // park the short tail at the front right after a launch: keeps the tuned kernels' 16-byte alignment and means
        // the buffer never has to slide while it holds a half-collected batch
        if (p->in.size() <= (1u << 16) && ((p->in.rd & 15) || p->in.rd > p->in.cap / 4)) SDR_TRY(p->in.realign(0));
    }
    return SDR_OK;
}

// copy n elements from host/device memory to dst on the ctx stream
static int fetch(sdr_pipe *p, void *d_dst, const void *src, size_t bytes, int mem) {
    if (!bytes) return SDR_OK;
    if (mem == SDR_HOST_PINNED) {
        if (p->pend_bytes && p->pend_src + p->pend_bytes == (const char *)src && p->pend_dst + p->pend_bytes == (char *)d_dst) {
            p->pend_bytes += bytes;   // extends the deferred copy
            // keep the DMA engine busy while the host is still collecting the rest of the batch
            if (p->pend_bytes >= ((size_t)8 << 20)) SDR_TRY(flush_pending(p));
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

static thread_local bool g_bound = false;   // sdr_pipe_run binds the device once for its whole loop

static int pipe_push_any(sdr_pipe *p, const void *src, long long n, int mem, long long n_vecs) {
    if (p->ps.open) {
        // fast path of an open persistent session: one more adjacent vector = one store the resident kernel will see
        const size_t bytes = (size_t)n * p->in_eb;
        if (mem == SDR_DEVICE_HELD && p->ext_p + p->ext_bytes == (const char *)src && n >= p->fir->T &&
            (long long)(p->ext_bytes + bytes) <= p->ps.capacity_bytes && !((PersistHdr *)p->ps.ctl)->error) {
            p->ext_bytes += bytes;
            p->n_total += n;
            ((PersistHdr *)p->ps.ctl)->published_bytes = (long long)p->ext_bytes;
            if (p->downstream) { persist_poll(p); return forward(p); }
            return SDR_OK;
        }
        if (!g_bound) SDR_TRY(p->ctx->bind());
        SDR_TRY(persist_close(p));
    }
    if (!g_bound) SDR_TRY(p->ctx->bind());
    TraceScope tr_push(trace_mode() == 1 ? "push(total)" : nullptr, p->ctx, n);
    if (is_fir_kind(p->kind)) {
        // the reference asserts every awaited vector holds at least numCoeffs samples (Filter.hs:544,586,691)
        long long need = is_resamp_kind(p->kind) ? (p->res->T + p->res->L - 1) / p->res->L : p->fir->T;
        if (is_byte_fed(p->kind)) {
            if (n & 1) return set_error(SDR_EINVAL, "u8 IQ stage: odd byte count %lld (interleaved I/Q pairs expected)", n);
            need *= 2;
        }
        if (n < need) return set_error(SDR_EPRECOND, "%s 1: input vector of %lld elements is shorter than numCoeffs (%lld)",
                                       p->assert_name, n, need);
        p->n_total += n;
        if (p->skip > 0) {   // samples a large decimation factor has already stepped over
            const long long d = p->skip < n ? p->skip : n;
            src = (const char *)src + (size_t)d * p->in_eb; n -= d; p->skip -= d;
            if (n == 0) return SDR_OK;
        }
        if (mem == SDR_DEVICE_HELD || mem == MEM_FWD) {
            // zero-copy: the vector is read in place by the next launch.  Vectors adjacent in memory extend the run; a
            // vector somewhere else first moves the run collected so far behind the stage's own buffer.
            SDR_TRY(flush_pending(p));
            const size_t bytes = (size_t)n * p->in_eb;
            if (p->ext_bytes && p->ext_fwd == (mem == MEM_FWD) && p->ext_p + p->ext_bytes == (const char *)src) p->ext_bytes += bytes;
            else { SDR_TRY(materialize_ext(p)); p->ext_p = (const char *)src; p->ext_bytes = bytes; }
            p->ext_fwd = (mem == MEM_FWD);   // what no launch consumes stays in the upstream stage's FIFO, held in place there
            SDR_TRY(process_fir(p));
            return forward(p);
        }
        SDR_TRY(materialize_ext(p));   // keeps the stream in order when held and copied pushes are mixed
        if (p->in.wr + (size_t)n * p->in_eb > p->in.cap) SDR_TRY(flush_pending(p));   // the buffer is about to move
        SDR_TRY(p->in.reserve((size_t)n * p->in_eb));
        SDR_TRY(fetch(p, p->in.p + p->in.wr, src, (size_t)n * p->in_eb, mem));
        p->in.wr += (size_t)n * p->in_eb;
        SDR_TRY(process_fir(p));
        return forward(p);
    }
    // element-wise kinds: one output vector per input vector (always read in place when the vector is on the device)
    if (mem == SDR_DEVICE_HELD || mem == MEM_FWD) mem = SDR_DEVICE;
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
    SDR_TRY(fifo_writable(p));
    SDR_TRY(fifo_reserve(p, (size_t)n_out * p->out_eb));
    void *d_dst = p->fifo.p + p->fifo.wr;
    if (p->kind == P_CONVERT) SDR_TRY(launch_convert_u8(p->ctx, (const uint8_t *)d_src, (float *)d_dst, n));
    else if (p->kind == P_SCALE) SDR_TRY(launch_scale(p->ctx, p->scale_k, (const float *)d_src, (float *)d_dst, n));
    else if (p->kind == P_DCBLOCK) SDR_TRY(launch_dc_blocker_carry(p->ctx, p->d_last, (const float *)d_src, (float *)d_dst, n));
    else {
        SDR_TRY(launch_fm_demod_carry(p->ctx, p->d_last, (const float *)d_src, (float *)d_dst, n));
        if (n) SDR_CUDA(cudaMemcpyAsync(p->d_last, (const char *)d_src + (size_t)(n - 1) * 8, 8, cudaMemcpyDeviceToDevice, p->ctx->stream));
    }
    p->fifo.wr += (size_t)n_out * p->out_eb;
    for (long long v = 0; v < n_vecs; v++) p->vec_lens.push_back(n_out / n_vecs);   // n_vecs equal vectors in one launch
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
int sdr_pipe_fm_frontend(sdr_decimator_t *d, int block_size_out, sdr_pipe_t **out) {
    if (!d || block_size_out <= 0) return set_error(SDR_EINVAL, "sdr_pipe_fm_frontend: bad argument");
    if (!d->r.cplx) return set_error(SDR_EINVAL, "sdr_pipe_fm_frontend: the decimator must be a complex-data one (fastDecimatorC)");
    sdr_pipe *p;
    SDR_TRY(new_pipe(d->r.ctx, P_FMFRONT, out, &p));
    p->fir = &d->r; p->in_eb = 1; p->out_eb = 4; p->block_out = block_size_out; p->assert_name = "decimate";
    p->bnd.c = p->scratch_x.c = p->scratch_y.c = p->ctx;
    SDR_CUDA(cudaMalloc(&p->d_last, 32));   // two carried samples (double buffer) + the fused kernel's ticket word
    SDR_CUDA(cudaMemsetAsync(p->d_last, 0, 32, p->ctx->stream));
    return SDR_OK;
}
int sdr_pipe_fm_lowrate(sdr_resampler_t *r, int block_size_resampler, sdr_filter_t *f, int block_size_out, float scale, sdr_pipe_t **out) {
    if (!r || !f || block_size_resampler <= 0 || block_size_out <= 0) return set_error(SDR_EINVAL, "sdr_pipe_fm_lowrate: bad argument");
    if (r->r.cplx || f->r.cplx) return set_error(SDR_EINVAL, "sdr_pipe_fm_lowrate: real-data resampler and filter expected (fm.hs:31-32)");
    if (r->r.ctx != f->r.ctx) return set_error(SDR_EINVAL, "sdr_pipe_fm_lowrate: records live on different contexts");
    if (block_size_resampler < f->r.T)
        return set_error(SDR_EPRECOND, "filter 1: upstream vectors of %d elements are shorter than numCoeffs (%d)", block_size_resampler, f->r.T);
    sdr_pipe *p;
    SDR_TRY(new_pipe(r->r.ctx, P_FMLOW, out, &p));
    p->res = &r->r; p->fir = &f->r; p->in_eb = p->out_eb = 4; p->block_out = block_size_out; p->block_r = block_size_resampler;
    p->scale_k = scale; p->assert_name = "resample";
    p->scratch_x.c = p->scratch_y.c = p->ctx;
    return SDR_OK;
}
int sdr_pipe_u8_decimator(sdr_decimator_t *d, int block_size_out, sdr_pipe_t **out) {
    if (!d || block_size_out <= 0) return set_error(SDR_EINVAL, "sdr_pipe_u8_decimator: bad argument");
    if (!d->r.cplx) return set_error(SDR_EINVAL, "sdr_pipe_u8_decimator: the decimator must be a complex-data one (fastDecimatorC)");
    sdr_pipe *p;
    SDR_TRY(new_pipe(d->r.ctx, P_U8DECIM, out, &p));
    p->fir = &d->r; p->in_eb = 1; p->out_eb = 8; p->block_out = block_size_out; p->assert_name = "decimate";
    p->scratch_x.c = p->ctx;
    return SDR_OK;
}
const char *sdr_pipe_last_kernel(const sdr_pipe_t *p) { return p ? p->last_kernel : "none"; }

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
int sdr_pipe_dc_blocker(sdr_ctx_t *ctx, sdr_pipe_t **out) {
    sdr_pipe *p;
    SDR_TRY(new_pipe(reinterpret_cast<Ctx *>(ctx), P_DCBLOCK, out, &p));
    p->in_eb = p->out_eb = 4;
    SDR_CUDA(cudaMalloc(&p->d_last, 8));   // (lastSample, lastOutput), both 0 at stream start (Filter.hs:731)
    SDR_CUDA(cudaMemsetAsync(p->d_last, 0, 8, p->ctx->stream));
    return SDR_OK;
}
int sdr_pipe_destroy(sdr_pipe_t *p) {
    if (!p) return SDR_OK;
    p->ctx->bind();
    cudaStreamSynchronize(p->ctx->stream);
    cudaStreamSynchronize(p->ctx->side);
    persist_close(p);
    if (p->ps.stream) cudaStreamDestroy(p->ps.stream);
    if (p->ps.ev) cudaEventDestroy(p->ps.ev);
    if (p->ps.ctl) cudaFreeHost(p->ps.ctl);
    if (p->ps.d_relay) cudaFree(p->ps.d_relay);
    // unlink: a neighbour that outlives this stage must not forward into (or be unlinked from) freed memory, nor keep
    // referring to vectors held in this stage's FIFO
    if (fifo_leased(p)) { materialize_ext(p->downstream); cudaStreamSynchronize(p->ctx->stream); }
    if (p->upstream) p->upstream->downstream = nullptr;
    if (p->downstream) p->downstream->upstream = nullptr;
    p->in.release(); p->fifo.release(); p->bnd.release(); p->scratch_x.release(); p->scratch_y.release();
    if (p->d_last) cudaFree(p->d_last);
    if (p->ev_out_ready) cudaEventDestroy(p->ev_out_ready);
    if (p->ev_out_done) cudaEventDestroy(p->ev_out_done);
    delete p;
    return SDR_OK;
}

int sdr_pipe_push(sdr_pipe_t *p, const void *in, long long n, int mem) {
    if (!p || n < 0 || (n && !in) || mem < SDR_HOST || mem > SDR_DEVICE_HELD)
        return set_error(SDR_EINVAL, "sdr_pipe_push: bad argument");
    if (n == 0) return SDR_OK;
    return pipe_push_any(p, in, n, mem);
}

int sdr_pipe_ready(sdr_pipe_t *p, int *n_blocks) {
    if (!p || !n_blocks) return set_error(SDR_EINVAL, "sdr_pipe_ready: bad argument");
    persist_poll(p);
    if (is_fir_kind(p->kind)) *n_blocks = (int)((long long)(p->fifo.size() / p->out_eb) / p->block_out);
    else *n_blocks = (int)p->vec_lens.size();
    return SDR_OK;
}

int sdr_pipe_next_len(sdr_pipe_t *p, long long *n) {
    if (!p || !n) return set_error(SDR_EINVAL, "sdr_pipe_next_len: bad argument");
    persist_poll(p);
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
    persist_poll(p);
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
    SDR_TRY(fifo_writable(p));
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
    if (p->ps.open) { SDR_TRY(persist_close(p)); SDR_TRY(process_fir(p, true)); SDR_TRY(forward(p)); }
    SDR_TRY(materialize_ext(p));   // SDR_DEVICE_HELD vectors are the caller's again after this call
    SDR_CUDA(cudaStreamSynchronize(p->ctx->stream));
    SDR_CUDA(cudaStreamSynchronize(p->ctx->side));
    trace_dump();
    return SDR_OK;
}

// Per-vector device pushes without a launch per vector: SDR_DEVICE_HELD vectors pushed back to back (adjacent in memory)
// are consumed by a RESIDENT kernel that polls how far the stream has been published and publishes the runs it has
// finished (kernels_fast.cu: k_dec_ring_persist).  max_session_samples bounds one session (the output FIFO is sized for
// it up front); 0 switches the mode off.  Complex decimators with up to 128 stored taps and decimation 8 only.
int sdr_pipe_set_persistent(sdr_pipe_t *p, long long max_session_samples) {
    if (!p || max_session_samples < 0) return set_error(SDR_EINVAL, "sdr_pipe_set_persistent: bad argument");
    int rs, hs;
    if (max_session_samples > 0 && !(p->kind == P_DECIM && p->fir->arith == SDR_ARITH_FAST &&
                                     dec_persist_geometry(p->fir->T, p->fir->D, p->fir->cplx, &rs, &hs)))
        return set_error(SDR_EINVAL, "sdr_pipe_set_persistent: no persistent consumer for this stage (complex decimate-by-8, up to 128 taps)");
    SDR_TRY(p->ctx->bind());
    SDR_TRY(persist_close(p));
    p->persist_max = max_session_samples;
    return SDR_OK;
}

int sdr_pipe_set_batch(sdr_pipe_t *p, long long min_outputs) {
    if (!p || min_outputs < 0) return set_error(SDR_EINVAL, "sdr_pipe_set_batch: bad argument");
    p->batch_min = min_outputs;
    if (is_fir_kind(p->kind) && min_outputs > 0) {
        // size both buffers for the batch once, instead of growing by doubling while the stream runs
        SDR_TRY(p->ctx->bind());
        SDR_TRY(persist_close(p));
        SDR_TRY(flush_pending(p));    // a reserve may slide or reallocate: no deferred copy may still target the old place,
        SDR_TRY(fifo_writable(p));    // and no in-flight drain may still be reading it
        long long in_per_out = is_resamp_kind(p->kind) ? (p->res->M + p->res->L - 1) / p->res->L
                             : is_byte_fed(p->kind) ? 2 * p->fir->D : p->fir->D;
        long long taps = is_resamp_kind(p->kind) ? p->res->T : p->fir->T;
        SDR_TRY(p->in.reserve((size_t)(2 * (min_outputs + p->block_out) * in_per_out + taps) * p->in_eb));
        SDR_TRY(fifo_reserve(p, (size_t)(2 * (min_outputs + p->block_out)) * p->out_eb));
    }
    return SDR_OK;
}

// ---- stream state export / import (SURVEY.md section 5 "checkpoint / resume", section 8b-ii) -------------------------------
// What a stage carries between vectors, and therefore what a checkpoint holds:
//   FIR kinds     the not-yet-consumed tail of the input stream (< numCoeffs samples plus what a batching threshold is
//                 still holding back) -- the reference's `crossover` carry (Filter.hs:558-569, 600-611, 712-727);
//   resampler     + the global output / input counters from which (group, offset) follow in closed form (Filter.hs:419-424);
//   fmDemod       the previous buffer's last sample (Demod.hs:41-46);  fused FM front end: the last decimated sample;
//   dcBlocker     (lastSample, lastOutput) (Filter.hs:731-739);
//   every kind    outputs produced but not yet popped (the partially filled output block of advanceOutBuf, Filter.hs:516-523).
namespace {
struct StateHeader {
    uint32_t magic, version;
    int32_t  kind, in_eb, out_eb, taps, factor, interp, last_sel, reserved;
    int64_t  block_out, k_next, pos, n_total, skip, in_bytes, fifo_bytes, n_vecs, last_bytes, block_r;
};
const uint32_t STATE_MAGIC = 0x50524453u;   // "SDRP"
size_t last_bytes_of(const sdr_pipe *p) {
    if (!p->d_last) return 0;
    return p->kind == P_FMFRONT ? 16 : 8;
}
void fill_header(const sdr_pipe *p, StateHeader *h) {
    memset(h, 0, sizeof(*h));
    h->magic = STATE_MAGIC; h->version = 1; h->kind = p->kind; h->in_eb = (int32_t)p->in_eb; h->out_eb = (int32_t)p->out_eb;
    if (p->fir) { h->taps = p->fir->T; h->factor = p->fir->D; h->interp = 1; }
    if (p->res) { h->taps = p->res->T; h->factor = p->res->M; h->interp = p->res->L; }
    if (p->kind == P_FMLOW) { h->reserved = p->fir->T; h->block_r = p->block_r; }
    h->last_sel = p->last_sel; h->block_out = p->block_out; h->k_next = p->k_next; h->pos = p->pos; h->n_total = p->n_total;
    h->skip = p->skip; h->in_bytes = is_fir_kind(p->kind) ? (int64_t)p->in.size() : 0; h->fifo_bytes = (int64_t)p->fifo.size();
    h->n_vecs = (int64_t)p->vec_lens.size(); h->last_bytes = (int64_t)last_bytes_of(p);
}
size_t state_bytes(const StateHeader &h) {
    return sizeof(StateHeader) + (size_t)h.in_bytes + (size_t)h.fifo_bytes + (size_t)h.n_vecs * 8 + (size_t)h.last_bytes;
}
}  // namespace

int sdr_pipe_state_size(sdr_pipe_t *p, size_t *bytes) {
    if (!p || !bytes) return set_error(SDR_EINVAL, "sdr_pipe_state_size: bad argument");
    StateHeader h;
    fill_header(p, &h);
    *bytes = state_bytes(h);
    return SDR_OK;
}

int sdr_pipe_state_save(sdr_pipe_t *p, void *buf, size_t capacity, size_t *written) {
    if (!p || !buf) return set_error(SDR_EINVAL, "sdr_pipe_state_save: bad argument");
    SDR_TRY(p->ctx->bind());
    SDR_TRY(flush_pending(p));
    SDR_TRY(persist_close(p));
    SDR_TRY(materialize_ext(p));
    SDR_TRY(fifo_writable(p));
    StateHeader h;
    fill_header(p, &h);
    const size_t need = state_bytes(h);
    if (capacity < need) return set_error(SDR_EINVAL, "sdr_pipe_state_save: %zu bytes needed, %zu given", need, capacity);
    char *q = (char *)buf;
    memcpy(q, &h, sizeof(h)); q += sizeof(h);
    cudaStream_t st = p->ctx->stream;
    if (h.in_bytes) SDR_CUDA(cudaMemcpyAsync(q, p->in.p + p->in.rd, (size_t)h.in_bytes, cudaMemcpyDeviceToHost, st));
    q += h.in_bytes;
    if (h.fifo_bytes) SDR_CUDA(cudaMemcpyAsync(q, p->fifo.p + p->fifo.rd, (size_t)h.fifo_bytes, cudaMemcpyDeviceToHost, st));
    q += h.fifo_bytes;
    for (long long n : p->vec_lens) { int64_t v = n; memcpy(q, &v, 8); q += 8; }
    if (h.last_bytes) SDR_CUDA(cudaMemcpyAsync(q, p->d_last, (size_t)h.last_bytes, cudaMemcpyDeviceToHost, st));
    SDR_CUDA(cudaStreamSynchronize(st));
    SDR_CUDA(cudaStreamSynchronize(p->ctx->side));
    if (written) *written = need;
    return SDR_OK;
}

// into a stage constructed the same way (same kind, taps / factors, block size); whatever it held is discarded
int sdr_pipe_state_restore(sdr_pipe_t *p, const void *buf, size_t bytes) {
    if (!p || !buf || bytes < sizeof(StateHeader)) return set_error(SDR_EINVAL, "sdr_pipe_state_restore: bad argument");
    StateHeader h, mine;
    memcpy(&h, buf, sizeof(h));
    fill_header(p, &mine);
    if (h.magic != STATE_MAGIC || h.version != 1) return set_error(SDR_EINVAL, "sdr_pipe_state_restore: not a pipe state (magic / version)");
    if (h.kind != mine.kind || h.in_eb != mine.in_eb || h.out_eb != mine.out_eb || h.taps != mine.taps || h.factor != mine.factor ||
        h.interp != mine.interp || h.block_out != mine.block_out || h.last_bytes != mine.last_bytes || h.reserved != mine.reserved ||
        h.block_r != mine.block_r)
        return set_error(SDR_EINVAL, "sdr_pipe_state_restore: the state was saved from a differently constructed stage "
                         "(kind %d/%d, taps %d/%d, factor %d/%d, block %lld/%lld)", h.kind, mine.kind, h.taps, mine.taps, h.factor,
                         mine.factor, (long long)h.block_out, (long long)mine.block_out);
    if (h.in_bytes < 0 || h.fifo_bytes < 0 || h.n_vecs < 0 || bytes < state_bytes(h))
        return set_error(SDR_EINVAL, "sdr_pipe_state_restore: truncated state (%zu bytes)", bytes);
    SDR_TRY(p->ctx->bind());
    SDR_TRY(flush_pending(p));
    SDR_TRY(fifo_writable(p));
    cudaStream_t st = p->ctx->stream;
    const char *q = (const char *)buf + sizeof(h);
    p->in.rd = p->in.wr = 0; p->fifo.rd = p->fifo.wr = 0; p->vec_lens.clear();
    p->ext_p = nullptr; p->ext_bytes = 0;
    if (h.in_bytes) {
        SDR_TRY(p->in.reserve((size_t)h.in_bytes));
        SDR_CUDA(cudaMemcpyAsync(p->in.p, q, (size_t)h.in_bytes, cudaMemcpyHostToDevice, st));
        p->in.wr = (size_t)h.in_bytes;
    }
    q += h.in_bytes;
    if (h.fifo_bytes) {
        SDR_TRY(fifo_reserve(p, (size_t)h.fifo_bytes));
        SDR_CUDA(cudaMemcpyAsync(p->fifo.p, q, (size_t)h.fifo_bytes, cudaMemcpyHostToDevice, st));
        p->fifo.wr = (size_t)h.fifo_bytes;
    }
    q += h.fifo_bytes;
    for (int64_t i = 0; i < h.n_vecs; i++) { int64_t v; memcpy(&v, q, 8); q += 8; p->vec_lens.push_back(v); }
    if (h.last_bytes) SDR_CUDA(cudaMemcpyAsync(p->d_last, q, (size_t)h.last_bytes, cudaMemcpyHostToDevice, st));
    p->last_sel = h.last_sel; p->k_next = h.k_next; p->pos = h.pos; p->n_total = h.n_total; p->skip = h.skip;
    SDR_CUDA(cudaStreamSynchronize(st));   // `buf` is the caller's again
    return SDR_OK;
}

// pop every complete vector of `sink` into out[written...] with ONE copy (FIR kinds) / one copy per vector otherwise
static int drain(sdr_pipe *sink, void *out, long long out_capacity, int out_mem, long long *written) {
    if (is_fir_kind(sink->kind)) {
        persist_poll(sink);
        long long nb = (long long)(sink->fifo.size() / sink->out_eb) / sink->block_out;
        if (nb == 0) return SDR_OK;
        // while a persistent consumer is resident, hand its outputs over in pieces of >= 32 vectors: a copy per vector would
        // cost more host time than the vector takes to compute (the session's end drains the rest)
        if (sink->ps.open && nb < 32) return SDR_OK;
        long long n = nb * sink->block_out;
        if (*written + n > out_capacity) return set_error(SDR_EINVAL, "sdr_pipe_run: output capacity %lld too small", out_capacity);
        TraceScope tr("drain", sink->ctx, n);
        if (out_mem == SDR_HOST_PINNED) {
            if (!sink->ev_out_ready) {
                SDR_CUDA(cudaEventCreateWithFlags(&sink->ev_out_ready, cudaEventDisableTiming));
                SDR_CUDA(cudaEventCreateWithFlags(&sink->ev_out_done, cudaEventDisableTiming));
            }
            SDR_TRY(fifo_writable(sink));   // at most one copy in flight per stage
            SDR_CUDA(cudaEventRecord(sink->ev_out_ready, sink->ctx->stream));
            SDR_CUDA(cudaStreamWaitEvent(sink->ctx->side, sink->ev_out_ready, 0));
            SDR_CUDA(cudaMemcpyAsync((char *)out + (size_t)*written * sink->out_eb, sink->fifo.p + sink->fifo.rd, (size_t)n * sink->out_eb,
                                     cudaMemcpyDeviceToHost, sink->ctx->side));
            SDR_CUDA(cudaEventRecord(sink->ev_out_done, sink->ctx->side));
            sink->d2h_outstanding = true;
        } else if (sink->fifo.borrowed && sink->fifo.p + sink->fifo.rd == (char *)out + (size_t)*written * sink->out_eb) {
            // produced in place: the vectors already are where the caller wants them
        } else {
            SDR_CUDA(cudaMemcpyAsync((char *)out + (size_t)*written * sink->out_eb, sink->fifo.p + sink->fifo.rd, (size_t)n * sink->out_eb,
                                     out_mem == SDR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, sink->ctx->stream));
            if (out_mem == SDR_HOST) SDR_CUDA(cudaStreamSynchronize(sink->ctx->stream));
        }
        sink->fifo.consume((size_t)n * sink->out_eb);
        *written += n;
        return SDR_OK;
    }
    // element-wise stage: the queued vectors are contiguous in the FIFO and are written back to back anyway
    long long total = 0;
    for (long long n : sink->vec_lens) total += n;
    if (total == 0) { sink->vec_lens.clear(); return SDR_OK; }
    if (*written + total > out_capacity) return set_error(SDR_EINVAL, "sdr_pipe_run: output capacity %lld too small", out_capacity);
    SDR_CUDA(cudaMemcpyAsync((char *)out + (size_t)*written * sink->out_eb, sink->fifo.p + sink->fifo.rd, (size_t)total * sink->out_eb,
                             out_mem == SDR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, sink->ctx->stream));
    if (out_mem == SDR_HOST) SDR_CUDA(cudaStreamSynchronize(sink->ctx->stream));
    sink->fifo.consume((size_t)total * sink->out_eb);
    sink->vec_lens.clear();
    *written += total;
    return SDR_OK;
}

// runEffect $ each vectors >-> p >-> ... >-> sink >-> collect, as one native loop
int sdr_pipe_run(sdr_pipe_t *p, sdr_pipe_t *sink, const void *in, long long vec_len, long long n_vecs, int in_mem,
                 void *out, long long out_capacity, int out_mem, long long *n_out) {
    if (!p || !sink || vec_len <= 0 || n_vecs < 0 || (n_vecs && !in) || (out_capacity && !out) || !n_out ||
        in_mem < SDR_HOST || in_mem > SDR_DEVICE_HELD || out_mem < SDR_HOST || out_mem > SDR_HOST_PINNED)
        return set_error(SDR_EINVAL, "sdr_pipe_run: bad argument");
    long long written = 0;
    SDR_TRY(p->ctx->bind());
    auto t_start = std::chrono::steady_clock::now();
    struct BoundGuard { BoundGuard() { g_bound = true; } ~BoundGuard() { g_bound = false; } } bound_guard;
    // on an error return no deferred or in-flight copy may still point into the caller's vectors (a no-op after the
    // final sdr_pipe_sync of the normal path)
    struct Quiesce {
        sdr_pipe *p; bool armed;
        ~Quiesce() { if (armed) { flush_pending(p); cudaStreamSynchronize(p->ctx->stream); cudaStreamSynchronize(p->ctx->side); } }
    } quiesce{p, true};
    // Device output: a FIR-kind sink produces straight into the caller's buffer (see fifo_return).  Whatever the stage
    // still holds from an earlier call (a partial output block) goes in front.  Given back on every exit path.
    struct Borrow {
        sdr_pipe *s;
        ~Borrow() { if (s && s->fifo.borrowed) { if (s->ps.open) persist_close(s); fifo_return(s); } }   // (error exits only)
    } borrow{nullptr};
    if (out_mem == SDR_DEVICE && is_fir_kind(sink->kind) && !sink->ps.open && !sink->downstream && !sink->fifo.borrowed &&
        (((uintptr_t)out) & 15) == 0 && sink->fifo.size() <= (size_t)out_capacity * sink->out_eb) {
        SDR_TRY(flush_pending(sink));
        SDR_TRY(fifo_writable(sink));
        const size_t live = sink->fifo.size();
        if (live) SDR_CUDA(cudaMemcpyAsync(out, sink->fifo.p + sink->fifo.rd, live, cudaMemcpyDeviceToDevice, sink->ctx->stream));
        sink->fifo_own = sink->fifo;
        LinBuf b;
        b.c = sink->ctx; b.p = (char *)out; b.cap = (size_t)out_capacity * sink->out_eb; b.rd = 0; b.wr = live; b.borrowed = true;
        sink->fifo = b;
        borrow.s = sink;
    }
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
    SDR_TRY(fifo_return(sink));
    auto t_issued = std::chrono::steady_clock::now();
    SDR_TRY(sdr_pipe_sync(p));
    if (trace_mode())
        fprintf(stderr, "[sdr_b200 trace] sdr_pipe_run: host issue %.1f us, final sync %.1f us\n",
                std::chrono::duration<double, std::micro>(t_issued - t_start).count(),
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_issued).count());
    quiesce.armed = false;   // sdr_pipe_sync above has done it
    *n_out = written;
    return SDR_OK;
}

// ---- producers / consumers on file descriptors ------------------------------------------------------------------------
namespace {

// one vector for fromHandle (Serialize.hs:82-83: PB.hGet blocks until `want` bytes or end of file) or for udpSource
// (NetworkStream.hs:33-35: one recv = one vector); returns bytes read, 0 at end of input, -1 on error
long long read_vector(int fd, char *dst, size_t want, bool datagram) {
    size_t got = 0;
    while (got < want) {
        ssize_t r = ::read(fd, dst + got, want - got);
        if (r < 0) { if (errno == EINTR) continue; return -1; }
        if (r == 0) break;
        got += (size_t)r;
        if (datagram) break;
    }
    return (long long)got;
}
bool write_all(int fd, const char *src, size_t bytes) {
    while (bytes) {
        ssize_t w = ::write(fd, src, bytes);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        src += w; bytes -= (size_t)w;
    }
    return true;
}
double seconds_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

struct FdRun {
    char *ring = nullptr, *obuf = nullptr;
    size_t half_bytes = 0, obuf_bytes = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    ~FdRun() {
        if (ring) cudaFreeHost(ring);
        if (obuf) cudaFreeHost(obuf);
        for (auto e : ev) if (e) cudaEventDestroy(e);
    }
};

// pop every complete vector of `sink` and write it to out_fd (toHandle, Serialize.hs:78-79; udpSink, NetworkStream.hs:37-42)
int drain_to_fd(sdr_pipe *sink, FdRun &R, int out_fd, bool datagram_out, sdr_io_stats_t *st) {
    std::vector<long long> lens;
    long long total = 0;
    if (is_fir_kind(sink->kind)) {
        long long nb = (long long)(sink->fifo.size() / sink->out_eb) / sink->block_out;
        for (long long i = 0; i < nb; i++) lens.push_back(sink->block_out);
        total = nb * sink->block_out;
    } else {
        for (long long n : sink->vec_lens) { lens.push_back(n); total += n; }
    }
    if (lens.empty()) return SDR_OK;
    const size_t bytes = (size_t)total * sink->out_eb;
    if (bytes > R.obuf_bytes) {
        if (R.obuf) { SDR_CUDA(cudaFreeHost(R.obuf)); R.obuf = nullptr; R.obuf_bytes = 0; }
        size_t cap = (size_t)1 << 20;
        while (cap < bytes) cap *= 2;
        SDR_CUDA(cudaHostAlloc((void **)&R.obuf, cap, cudaHostAllocDefault));
        R.obuf_bytes = cap;
    }
    long long written = 0;
    SDR_TRY(drain(sink, R.obuf, (long long)(R.obuf_bytes / sink->out_eb), SDR_HOST_PINNED, &written));
    // the copy rides the side stream (FIR kinds) or the main stream (element-wise kinds): wait for it, not for the
    // launches queued behind it
    if (sink->d2h_outstanding) SDR_CUDA(cudaEventSynchronize(sink->ev_out_done));
    else SDR_CUDA(cudaStreamSynchronize(sink->ctx->stream));
    st->vectors_out += (long long)lens.size();
    st->elements_out += written;
    if (out_fd < 0) return SDR_OK;
    auto t0 = std::chrono::steady_clock::now();
    bool ok = true;
    if (datagram_out) {
        const char *q = R.obuf;
        for (long long n : lens) { ok = ok && write_all(out_fd, q, (size_t)n * sink->out_eb); q += (size_t)n * sink->out_eb; }
    } else {
        ok = write_all(out_fd, R.obuf, (size_t)written * sink->out_eb);
    }
    st->write_seconds += seconds_since(t0);
    if (!ok) return set_error(SDR_EINVAL, "sdr_pipe_run_fd: write to descriptor %d failed: %s", out_fd, strerror(errno));
    return SDR_OK;
}

}  // namespace

int sdr_pipe_run_fd(sdr_pipe_t *p, sdr_pipe_t *sink, int in_fd, long long vec_len, long long max_vecs, int out_fd, int flags,
                    sdr_io_stats_t *stats) {
    if (!p || !sink || in_fd < 0 || vec_len <= 0 || max_vecs < 0)
        return set_error(SDR_EINVAL, "sdr_pipe_run_fd: bad argument");
    const bool dgram_in = flags & SDR_IO_DATAGRAM_IN, dgram_out = flags & SDR_IO_DATAGRAM_OUT;
    if (dgram_in && max_vecs == 0)
        return set_error(SDR_EINVAL, "sdr_pipe_run_fd: a datagram source never ends, max_vecs must be given");
    sdr_io_stats_t local = {0, 0, 0, 0, 0.0, 0.0};
    sdr_io_stats_t *st = stats ? stats : &local;
    *st = local;
    SDR_TRY(p->ctx->bind());
    struct BoundGuard { BoundGuard() { g_bound = true; } ~BoundGuard() { g_bound = false; } } bound_guard;
    FdRun R;
    // whatever way this function is left, no deferred or in-flight copy may still point into R's page-locked ring when
    // it is freed (declared after R: runs before R's destructor)
    struct Quiesce {
        sdr_pipe *p;
        ~Quiesce() { flush_pending(p); cudaStreamSynchronize(p->ctx->stream); cudaStreamSynchronize(p->ctx->side); }
    } quiesce{p};
    const size_t vec_bytes = (size_t)vec_len * p->in_eb;
    // page-locked staging ring of two halves: read() lands directly in DMA-able memory; while one half is being
    // copied to the device the other one is being filled
    size_t per_half = ((size_t)8 << 20) / vec_bytes;
    if (per_half < 1) per_half = 1;
    R.half_bytes = per_half * vec_bytes;
    SDR_CUDA(cudaHostAlloc((void **)&R.ring, 2 * R.half_bytes, cudaHostAllocDefault));
    for (auto &e : R.ev) SDR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
    bool used[2] = {false, false};
    int half = 0;
    size_t fill = 0;
    int deferred = SDR_OK;
    for (long long v = 0; max_vecs == 0 || v < max_vecs; v++) {
        if (fill + vec_bytes > R.half_bytes) {
            SDR_TRY(flush_pending(p));
            SDR_CUDA(cudaEventRecord(R.ev[half], p->ctx->stream));
            used[half] = true;
            half ^= 1; fill = 0;
            if (used[half]) SDR_CUDA(cudaEventSynchronize(R.ev[half]));   // its copies to the device have finished
        }
        char *dst = R.ring + (size_t)half * R.half_bytes + fill;
        auto t0 = std::chrono::steady_clock::now();
        long long got = read_vector(in_fd, dst, vec_bytes, dgram_in);
        st->read_seconds += seconds_since(t0);
        if (got < 0) return set_error(SDR_EINVAL, "sdr_pipe_run_fd: read from descriptor %d failed: %s", in_fd, strerror(errno));
        if (got == 0) break;
        if ((size_t)got % p->in_eb)
            return set_error(SDR_EINVAL, "sdr_pipe_run_fd: %lld bytes read are not a whole number of %zu-byte elements", got, p->in_eb);
        const long long n = got / (long long)p->in_eb;
        int s = pipe_push_any(p, dst, n, SDR_HOST_PINNED);
        if (s == SDR_EPRECOND && (size_t)got < vec_bytes && !dgram_in) { deferred = s; break; }   // short last vector: the reference's assert
        SDR_TRY(s);
        fill += (size_t)got;
        st->vectors_in++; st->elements_in += n;
        SDR_TRY(drain_to_fd(sink, R, out_fd, dgram_out, st));
        if ((size_t)got < vec_bytes && !dgram_in) break;   // end of file inside the vector
    }
    for (sdr_pipe *q = p; q; q = q->downstream) {
        if (is_fir_kind(q->kind)) { SDR_TRY(process_fir(q, true)); SDR_TRY(forward(q)); }
        if (q == sink) break;
    }
    SDR_TRY(drain_to_fd(sink, R, out_fd, dgram_out, st));
    SDR_TRY(sdr_pipe_sync(p));
    return deferred;
}

int sdr_pipe_connect(sdr_pipe_t *src, sdr_pipe_t *dst) {
    if (!src || !dst || src == dst) return set_error(SDR_EINVAL, "sdr_pipe_connect: bad argument");
    if (src->ctx != dst->ctx) return set_error(SDR_EINVAL, "sdr_pipe_connect: stages live on different contexts");
    if (src->out_eb != dst->in_eb && !(dst->kind == P_CONVERT))
        return set_error(SDR_EINVAL, "sdr_pipe_connect: element types differ (%zu-byte out, %zu-byte in)", src->out_eb, dst->in_eb);
    if (is_fir_kind(src->kind) && is_fir_kind(dst->kind)) {
        long long need = is_resamp_kind(dst->kind) ? (dst->res->T + dst->res->L - 1) / dst->res->L : dst->fir->T;
        if (src->block_out < need)
            return set_error(SDR_EPRECOND, "%s 1: upstream vectors of %lld elements are shorter than numCoeffs (%lld)",
                             dst->assert_name, src->block_out, need);
    }
    if (dst->upstream && dst->upstream != src)
        return set_error(SDR_EINVAL, "sdr_pipe_connect: the destination stage already has an upstream stage");
    if (src->downstream && src->downstream != dst) {
        if (fifo_leased(src)) { SDR_TRY(src->ctx->bind()); SDR_TRY(materialize_ext(src->downstream)); }
        src->downstream->upstream = nullptr;
    }
    src->downstream = dst;
    dst->upstream = src;
    return SDR_OK;
}

}  // extern "C"
