// multigpu.cu -- sharding a flat stream over the GPUs of one box (the reference is single
// process and has no counterpart -- the halo is what its `crossover` state carries between buffers,
// hs_sources/SDR/Filter.hs:600-611), plus the synthetic-stream / checksum / L2-flush helpers of the C ABI.
//
// Each rank keeps one contiguous chunk of the input resident and owns the outputs whose windows START in it.  The
// windows of its last (T - D) / D outputs run into the next rank's chunk: those T - D samples are the only data
// exchanged (ncclSend to rank-1 / ncclRecv from rank+1 on a side stream) while the interior outputs are computed;
// the boundary outputs follow once the halo has landed.  NCCL is bound at run time with dlopen so the library
// loads (and every single-GPU path works) without it and shares the copy of libnccl a host process already has.
#include "records.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <vector>

namespace sdr {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static int nccl_api(NcclApi **out) {
    static NcclApi api;
    static int state = 0;   // 0 untried, 1 ok, -1 failed
    if (state == 0) {
        const char *names[] = {getenv("SDR_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        state = -1;
        if (api.lib) {
#define SDR_SYM(field, name) *(void **)(&api.field) = dlsym(api.lib, name)
            SDR_SYM(GetUniqueId, "ncclGetUniqueId");
            SDR_SYM(CommInitRank, "ncclCommInitRank");
            SDR_SYM(CommDestroy, "ncclCommDestroy");
            SDR_SYM(GroupStart, "ncclGroupStart");
            SDR_SYM(GroupEnd, "ncclGroupEnd");
            SDR_SYM(Send, "ncclSend");
            SDR_SYM(Recv, "ncclRecv");
            SDR_SYM(AllGather, "ncclAllGather");
            SDR_SYM(AllReduce, "ncclAllReduce");
            SDR_SYM(GetErrorString, "ncclGetErrorString");
#undef SDR_SYM
            if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send &&
                api.Recv && api.AllGather && api.AllReduce && api.GetErrorString)
                state = 1;
        }
    }
    if (state != 1) return set_error(SDR_ENCCL, "NCCL not available: %s", api.lib ? "missing symbols" : dlerror());
    *out = &api;
    return SDR_OK;
}

#define SDR_NCCL(api, expr)                                                                       \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess) return set_error(SDR_ENCCL, "NCCL error %d (%s) in %s", (int)_r,   \
                                                (api)->GetErrorString(_r), #expr);               \
    } while (0)

}  // namespace sdr

using namespace sdr;

struct sdr_comm {
    Ctx *ctx = nullptr;
    NcclApi *api = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    void *d_halo = nullptr;
    size_t halo_bytes = 0;
    // peer transport: the right neighbour's registered chunk, mapped into this process with CUDA IPC
    const void *my_base = nullptr;   // the allocation this rank registered
    void *peer_base = nullptr;       // rank+1's allocation as seen from here (nullptr: NCCL transport)
    cudaEvent_t ev_ready = nullptr, ev_halo = nullptr;
    int *d_token = nullptr;          // sdr_comm_barrier's all-reduce operand
};

extern "C" {

int sdr_shard_plan(long long n_samples, int taps, int factor, int world, int rank, sdr_shard_t *plan) {
    if (!plan || n_samples < 0 || taps <= 0 || factor <= 0 || world <= 0 || rank < 0 || rank >= world)
        return set_error(SDR_EINVAL, "sdr_shard_plan: bad argument");
    long long total_out = n_samples >= taps ? (n_samples - taps) / factor + 1 : 0;
    // chunk: ceil(N / world) rounded up to whole 256-output tiles so every rank's chunk starts tile- and 16-byte aligned
    long long unit = 256LL * factor;
    long long chunk = (n_samples + world - 1) / world;
    chunk = ((chunk + unit - 1) / unit) * unit;
    long long in_begin = chunk * rank;
    if (in_begin > n_samples) in_begin = n_samples;
    long long in_end = in_begin + chunk;
    if (in_end > n_samples) in_end = n_samples;
    long long out_begin = (in_begin + factor - 1) / factor;          // first window starting inside the chunk
    long long out_end = (in_end + factor - 1) / factor;
    if (out_end > total_out) out_end = total_out;
    if (out_begin > out_end) out_begin = out_end;
    plan->in_begin = in_begin;
    plan->in_count = in_end - in_begin;
    plan->out_begin = out_begin;
    plan->out_count = out_end - out_begin;
    long long last_needed = plan->out_count ? (out_end - 1) * factor + taps : in_end;   // one past the last sample read
    plan->halo = last_needed > in_end ? last_needed - in_end : 0;
    long long local0 = out_begin * factor - in_begin;                // offset of the first owned window in the chunk
    long long interior = (plan->in_count - local0 >= taps) ? (plan->in_count - local0 - taps) / factor + 1 : 0;
    plan->out_interior = interior < plan->out_count ? interior : plan->out_count;
    plan->n_samples = n_samples; plan->taps = taps; plan->factor = factor; plan->world = world; plan->rank = rank;
    return SDR_OK;
}

int sdr_comm_unique_id(unsigned char id[SDR_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == SDR_COMM_ID_BYTES, "ncclUniqueId size");
    NcclApi *api;
    SDR_TRY(nccl_api(&api));
    ncclUniqueId u;
    SDR_NCCL(api, api->GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return SDR_OK;
}

int sdr_comm_create(sdr_ctx_t *ctx, const unsigned char id[SDR_COMM_ID_BYTES], int world, int rank, sdr_comm_t **out) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !id || !out || world <= 0 || rank < 0 || rank >= world) return set_error(SDR_EINVAL, "sdr_comm_create: bad argument");
    *out = nullptr;
    NcclApi *api;
    SDR_TRY(nccl_api(&api));
    SDR_TRY(c->bind());
    sdr_comm *h = new sdr_comm();
    h->ctx = c; h->api = api; h->world = world; h->rank = rank;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = api->CommInitRank(&h->comm, world, u, rank);
    if (r != ncclSuccess) { delete h; return set_error(SDR_ENCCL, "ncclCommInitRank failed: %s", api->GetErrorString(r)); }
    cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming);
    *out = h;
    return SDR_OK;
}

int sdr_comm_destroy(sdr_comm_t *c) {
    if (!c) return SDR_OK;
    c->ctx->bind();
    cudaStreamSynchronize(c->ctx->stream);
    cudaStreamSynchronize(c->ctx->side);
    if (c->comm) c->api->CommDestroy(c->comm);
    if (c->d_halo) cudaFree(c->d_halo);
    if (c->d_token) cudaFree(c->d_token);
    if (c->peer_base) cudaIpcCloseMemHandle(c->peer_base);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_halo) cudaEventDestroy(c->ev_halo);
    delete c;
    return SDR_OK;
}

// Collective.  Every rank passes the BASE address of the device allocation that holds its chunk (as returned by
// sdr_dev_alloc).  The CUDA IPC handles travel by ncclAllGather; each rank maps its right neighbour's allocation.
// Afterwards sdr_decimate_sharded on that same d_in reads the T-D halo samples straight out of the neighbour's HBM over
// NVLink inside the boundary launch: no NCCL kernel per pass, no rendezvous, no SMs set aside.  The caller guarantees
// that the neighbour's chunk is complete before a pass starts (e.g. a barrier after producing it).
int sdr_comm_share_chunks(sdr_comm_t *c, const void *d_chunk_base) {
    if (!c) return set_error(SDR_EINVAL, "sdr_comm_share_chunks: bad argument");
    Ctx *ctx = c->ctx;
    SDR_TRY(ctx->bind());
    if (c->peer_base) {
        SDR_CUDA(cudaStreamSynchronize(ctx->stream));   // no pass may still be reading through the mapping
        cudaIpcCloseMemHandle(c->peer_base); c->peer_base = nullptr;
    }
    c->my_base = nullptr;
    if (!d_chunk_base) return SDR_OK;   // NULL: drop the mapping, back to the NCCL transport (local, not a collective)
    if (c->world == 1) return SDR_OK;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t mine;
    SDR_CUDA(cudaIpcGetMemHandle(&mine, const_cast<void *>(d_chunk_base)));
    unsigned char *d_all = nullptr;
    SDR_CUDA(cudaMalloc(&d_all, 64 * (size_t)(c->world + 1)));
    SDR_CUDA(cudaMemcpyAsync(d_all + 64 * (size_t)c->world, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t r = c->api->AllGather(d_all + 64 * (size_t)c->world, d_all, 64, ncclChar, c->comm, ctx->stream);
    if (r != ncclSuccess) { cudaFree(d_all); return set_error(SDR_ENCCL, "ncclAllGather of IPC handles failed: %s", c->api->GetErrorString(r)); }
    std::vector<cudaIpcMemHandle_t> all(c->world);
    SDR_CUDA(cudaMemcpyAsync(all.data(), d_all, 64 * (size_t)c->world, cudaMemcpyDeviceToHost, ctx->stream));
    SDR_CUDA(cudaStreamSynchronize(ctx->stream));
    SDR_CUDA(cudaFree(d_all));
    if (c->rank + 1 < c->world) {
        void *peer = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&peer, all[c->rank + 1], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle(right neighbour)", __FILE__, __LINE__);
        c->peer_base = peer;
    }
    c->my_base = d_chunk_base;
    return SDR_OK;
}

// In-stream rendezvous: a 4-byte ncclAllReduce on the ctx stream.  Work enqueued behind it starts on every rank within
// the collective's completion skew (microseconds) -- unlike a host-side barrier, which leaves each rank's launch latency
// and host jitter inside whatever is timed next.
int sdr_comm_barrier(sdr_comm_t *c) {
    if (!c) return set_error(SDR_EINVAL, "sdr_comm_barrier: null communicator");
    Ctx *ctx = c->ctx;
    SDR_TRY(ctx->bind());
    if (c->world == 1) return SDR_OK;
    if (!c->d_token) { SDR_CUDA(cudaMalloc(&c->d_token, 8)); SDR_CUDA(cudaMemsetAsync(c->d_token, 0, 8, ctx->stream)); }
    SDR_NCCL(c->api, c->api->AllReduce(c->d_token, c->d_token + 1, 1, ncclInt, ncclSum, c->comm, ctx->stream));
    return SDR_OK;
}

// 1 when sdr_decimate_sharded on this d_in would use the peer-memory transport
int sdr_comm_peer_halo_active(const sdr_comm_t *c, const void *d_in) {
    return c && c->my_base && c->my_base == d_in ? 1 : 0;
}

int sdr_decimate_sharded(sdr_decimator_t *d, sdr_comm_t *c, const sdr_shard_t *plan, const void *d_in, void *d_out) {
    if (!d || !plan || (plan->in_count && !d_in) || (plan->out_count && !d_out))
        return set_error(SDR_EINVAL, "sdr_decimate_sharded: bad argument");
    FirRec &r = d->r;
    Ctx *ctx = r.ctx;
    if (plan->taps != r.T || plan->factor != r.D)
        return set_error(SDR_EINVAL, "sdr_decimate_sharded: plan was made for %d taps / %d, decimator has %d / %d", plan->taps,
                         plan->factor, r.T, r.D);
    if (plan->world > 1 && (!c || c->world != plan->world || c->rank != plan->rank || c->ctx != ctx))
        return set_error(SDR_EINVAL, "sdr_decimate_sharded: communicator does not match the plan");
    SDR_TRY(ctx->bind());
    const size_t eb = elem_bytes(r.cplx);
    const long long local0 = plan->out_begin * plan->factor - plan->in_begin;
    bool exchange = false;
    const bool peer = plan->world > 1 && c->my_base && c->my_base == d_in;
    // the windows of this rank's last outputs run `halo` samples into the right neighbour's chunk: that chunk must hold
    // them (evaluated identically on every rank from the plan alone, so all ranks fail together -- none is left blocked
    // in a collective)
    for (int q = 0; q + 1 < plan->world; q++) {
        sdr_shard_t a, b;
        SDR_TRY(sdr_shard_plan(plan->n_samples, plan->taps, plan->factor, plan->world, q, &a));
        SDR_TRY(sdr_shard_plan(plan->n_samples, plan->taps, plan->factor, plan->world, q + 1, &b));
        if (a.halo > b.in_count)
            return set_error(SDR_EPRECOND, "sdr_decimate_sharded: rank %d's chunk of %lld samples is shorter than the %lld-sample halo of rank %d",
                             q + 1, b.in_count, a.halo, q);
    }
    if (peer) {
        // Halo read in place from the neighbour's HBM over NVLink, INSIDE the ring kernel: its edge fills take the last
        // T-D samples of the boundary windows from the second source pointer (TMA bulk copies of peer memory), so a pass
        // is ONE launch -- no NCCL kernel, no rendezvous, no side stream, no SMs set aside.
        if (plan->halo > 0 && !c->peer_base) return set_error(SDR_EPRECOND, "sdr_decimate_sharded: no mapped right neighbour");
        Seg2 seg = {d_in, plan->in_count, plan->halo > 0 ? c->peer_base : nullptr, plan->halo};
        return r.run(seg, local0, d_out, plan->out_count, false);
    }
    if (plan->world > 1) {
        // what my left neighbour needs from the head of my chunk
        sdr_shard_t left;
        long long send = 0;
        if (plan->rank > 0) { SDR_TRY(sdr_shard_plan(plan->n_samples, plan->taps, plan->factor, plan->world, plan->rank - 1, &left));
                              send = left.halo; }
        long long recv = plan->halo;
        if (send || recv) {
            exchange = true;
            if ((size_t)recv * eb > c->halo_bytes) {
                if (c->d_halo) { SDR_CUDA(cudaStreamSynchronize(ctx->side)); SDR_CUDA(cudaFree(c->d_halo)); }
                c->halo_bytes = (size_t)recv * eb + 256;
                SDR_CUDA(cudaMalloc(&c->d_halo, c->halo_bytes));
            }
            // the chunk was produced on the main stream: the side stream must not send it early
            SDR_CUDA(cudaEventRecord(c->ev_ready, ctx->stream));
            SDR_CUDA(cudaStreamWaitEvent(ctx->side, c->ev_ready, 0));
            SDR_NCCL(c->api, c->api->GroupStart());
            if (send) SDR_NCCL(c->api, c->api->Send(d_in, (size_t)send * eb, ncclChar, plan->rank - 1, c->comm, ctx->side));
            if (recv) SDR_NCCL(c->api, c->api->Recv(c->d_halo, (size_t)recv * eb, ncclChar, plan->rank + 1, c->comm, ctx->side));
            SDR_NCCL(c->api, c->api->GroupEnd());
        }
    }
    const char *kernel = r.last_kernel;
    if (!exchange) {   // single rank (or nothing to exchange): plain stream pass
        if (plan->out_count > plan->out_interior) return set_error(SDR_EPRECOND, "sdr_decimate_sharded: boundary outputs without a neighbour");
        Seg2 seg = {d_in, plan->in_count, nullptr, 0};
        return r.run(seg, local0, d_out, plan->out_count, false);
    }
    // main stream: the tuned kernel over every sub-tile that lies inside the resident chunk.  The NCCL send/recv kernel
    // cannot share an SM with a CTA of the persistent decimator (registers): leave it a few SMs, otherwise the CTAs
    // queued behind it start only after the rendezvous and stretch the whole pass.
    long long done = 0;
    ctx->reserve_sms = 4;
    int rc = r.run_tuned(d_in, plan->in_count, local0, d_out, plan->out_interior, &done);
    ctx->reserve_sms = 0;
    SDR_TRY(rc);
    if (done > 0) kernel = r.last_kernel;
    // side stream, right behind the receive: ONE generic launch for the ragged end of the interior plus the windows
    // that run into the halo; the main stream joins once at the end.
    if (plan->out_count > done) {
        Seg2 seg = {d_in, plan->in_count, c->d_halo, plan->halo};
        ctx->override_st = ctx->side;
        rc = r.run(seg, local0 + done * plan->factor, (char *)d_out + (size_t)done * eb, plan->out_count - done, false);
        ctx->override_st = nullptr;
        SDR_TRY(rc);
    }
    SDR_CUDA(cudaEventRecord(c->ev_halo, ctx->side));
    SDR_CUDA(cudaStreamWaitEvent(ctx->stream, c->ev_halo, 0));   // also keeps the send ordered before later writes to d_in
    r.last_kernel = kernel;
    return SDR_OK;
}

// ---- element-wise kernels on device-resident buffers (enqueue-only; the one-shot layer-1 symbols stage host memory) ----
int sdr_dev_convert_u8(sdr_ctx_t *ctx, const uint8_t *d_in, float *d_out, long long n_bytes) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || n_bytes < 0 || (n_bytes && (!d_in || !d_out))) return set_error(SDR_EINVAL, "sdr_dev_convert_u8: bad argument");
    return launch_convert_u8(c, d_in, d_out, n_bytes);
}
int sdr_dev_scale(sdr_ctx_t *ctx, float factor, const float *d_in, float *d_out, long long n) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || n < 0 || (n && (!d_in || !d_out))) return set_error(SDR_EINVAL, "sdr_dev_scale: bad argument");
    return launch_scale(c, factor, d_in, d_out, n);
}
int sdr_dev_fm_demod(sdr_ctx_t *ctx, float last_re, float last_im, const float *d_in, float *d_out, long long n) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || n < 0 || (n && (!d_in || !d_out))) return set_error(SDR_EINVAL, "sdr_dev_fm_demod: bad argument");
    return launch_fm_demod(c, last_re, last_im, d_in, d_out, n);
}

// ---- synthetic streams + measurement helpers -----------------------------------------------------------------------
int sdr_synth_noise(sdr_ctx_t *ctx, float *d_out, long long n_floats, long long first_float, uint32_t seed) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || n_floats < 0 || (n_floats && !d_out)) return set_error(SDR_EINVAL, "sdr_synth_noise: bad argument");
    return launch_synth_noise(c, d_out, n_floats, first_float, seed);
}
int sdr_synth_bytes(sdr_ctx_t *ctx, uint8_t *d_out, long long n_bytes, long long first_byte, uint32_t seed) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || n_bytes < 0 || (n_bytes && !d_out)) return set_error(SDR_EINVAL, "sdr_synth_bytes: bad argument");
    return launch_synth_bytes(c, d_out, n_bytes, first_byte, seed);
}
int sdr_flush_l2(sdr_ctx_t *ctx) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return set_error(SDR_EINVAL, "sdr_flush_l2: null ctx");
    SDR_TRY(c->bind());
    if (!c->d_flush) { c->d_flush_bytes = (size_t)256 << 20; SDR_CUDA(cudaMalloc(&c->d_flush, c->d_flush_bytes)); }
    return launch_fill(c, c->d_flush, c->d_flush_bytes);
}
int sdr_checksum32(sdr_ctx_t *ctx, const void *d_buf, long long n_words, long long first_word, uint64_t *sum) {
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !sum || n_words < 0 || (n_words && !d_buf)) return set_error(SDR_EINVAL, "sdr_checksum32: bad argument");
    SDR_TRY(c->bind());
    SDR_TRY(c->ensure_stage(0, 64));
    unsigned long long *d_sum = (unsigned long long *)c->d_stage_out;
    SDR_TRY(launch_checksum32(c, (const uint32_t *)d_buf, n_words, first_word, d_sum));
    unsigned long long h = 0;
    SDR_CUDA(cudaMemcpyAsync(&h, d_sum, 8, cudaMemcpyDeviceToHost, c->stream));
    SDR_CUDA(cudaStreamSynchronize(c->stream));
    *sum = h;
    return SDR_OK;
}

}  // extern "C"
