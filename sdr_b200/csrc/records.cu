// records.cu -- layer 1 (reference-signature one-shot entry points) and layer 2 (plugin records) of the C ABI.
//
// Layer 1 replaces the C symbols the reference imports with `foreign import ccall unsafe`
// (hs_sources/SDR/FilterInternal.hs:80-249, 346-388; hs_sources/SDR/Util.hs:100-241): same argument order and
// meaning, HOST pointers, blocking.  Layer 2 replaces the records of closures `Filter` / `Decimator` / `Resampler`
// (hs_sources/SDR/Filter.hs:116-144) and their constructors (:163-502).
#include "records.cuh"

#include <cstdlib>
#include <cstring>

namespace sdr {

// launchers defined in kernels_generic.cu that are not in common.cuh
int launch_dc_blocker(Ctx *c, float last_sample, float last_output, const float *d_in, float *d_out, long long n,
                      float *d_final2);

static int round_up(int num, int div) { return ((num + div - 1) / div) * div; }   // FilterInternal.hs:287-288

// ---------------------------------------------------------------------------------------------------------------
// Staged
// ---------------------------------------------------------------------------------------------------------------
int Staged::begin() {
    SDR_TRY(c->bind());
    if (mem == SDR_DEVICE) { d_in = in; d_out = out; return SDR_OK; }
    SDR_TRY(c->ensure_stage(in_bytes, out_bytes));
    if (in_bytes) SDR_CUDA(cudaMemcpyAsync(c->d_stage_in, in, in_bytes, cudaMemcpyHostToDevice, c->stream));
    d_in = c->d_stage_in;
    d_out = c->d_stage_out;
    return SDR_OK;
}
int Staged::end() {
    if (mem == SDR_DEVICE) return SDR_OK;
    if (out_bytes) SDR_CUDA(cudaMemcpyAsync(out, c->d_stage_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    SDR_CUDA(cudaStreamSynchronize(c->stream));
    return SDR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// FirRec
// ---------------------------------------------------------------------------------------------------------------
int FirRec::create(Ctx *c, bool is_complex, int factor, const float *coeffs, int n, int size_multiple, bool sym_half) {
    if (!c || !coeffs || n <= 0 || factor <= 0 || size_multiple <= 0)
        return set_error(SDR_EINVAL, "FIR constructor: bad argument (taps %d, factor %d, multiple %d)", n, factor, size_multiple);
    ctx = c; cplx = is_complex; D = factor; arith = c->arith;
    SDR_TRY(c->bind());
    std::vector<float> full;
    if (sym_half) {   // mkFilterSymR / mkDecimatorSymR Filter.hs:234-245, 358-371: numCoeffs = 2 * half, no padding
        full.assign(coeffs, coeffs + n);
        for (int i = n - 1; i >= 0; i--) full.push_back(coeffs[i]);
        T = 2 * n;
    } else {          // mkFilter / mkDecimator Filter.hs:163-175, 277-290: zero pad to the size multiple
        T = round_up(n, size_multiple);
        full.assign(coeffs, coeffs + n);
        full.resize(T, 0.0f);
    }
    symmetric = true;
    for (int k = 0; k < T / 2; k++) if (memcmp(&full[k], &full[T - 1 - k], sizeof(float)) != 0) { symmetric = false; break; }
    // the tuned kernels are instantiated for 32 / 64 / 128 taps and read their whole capacity: zero-padded on the device
    h_taps.assign(full.begin(), full.begin() + T);
    const int t_alloc = round_up(T, 128) + 128;
    full.resize(t_alloc, 0.0f);
    SDR_CUDA(cudaMalloc(&d_taps, sizeof(float) * t_alloc));
    SDR_CUDA(cudaMemcpyAsync(d_taps, full.data(), sizeof(float) * t_alloc, cudaMemcpyHostToDevice, c->stream));
    // EXACT: the AVX member of the family (CPUID.hs:100-104 picks AVX first)
    if (sym_half) { ex_W = 8; ex_layout = cplx ? 2 : 0; ex_sym = 1; ex_T = n; }
    else          { ex_W = 8; ex_layout = cplx ? 1 : 0; ex_sym = 0; ex_T = T; }
    d_ex_taps = d_taps;   // the first ex_T entries are exactly what the exact kernel indexes (half or plain taps)
    SDR_CUDA(cudaStreamSynchronize(c->stream));   // `full` dies with this frame
    return SDR_OK;
}

void FirRec::destroy() {
    if (ctx && d_taps) { ctx->bind(); cudaFree(d_taps); }
    d_taps = d_ex_taps = nullptr;
}

int FirRec::run(Seg2 seg, long long first, void *d_out, long long num, bool cross_order) {
    if (num <= 0) return SDR_OK;
    const size_t eb = elem_bytes(cplx);
    // fold `first` into the segments
    if (first >= seg.na) { seg.a = (const char *)seg.b + (first - seg.na) * eb; seg.na = seg.nb - (first - seg.na);
                           seg.b = nullptr; seg.nb = 0; if (seg.na < 0) seg.na = 0; }
    else                 { seg.a = (const char *)seg.a + first * eb; seg.na -= first; }
    if (arith == SDR_ARITH_EXACT) {
        last_kernel = "fir_exact";
        if (cross_order)   // the cross-buffer kernels are strict left-to-right sums (FilterInternal.hs:398-408)
            return launch_fir_exact_fir(ctx, cplx, T, D, 1, 0, 0, d_taps, seg, d_out, num);
        return launch_fir_exact_fir(ctx, cplx, ex_T, D, ex_W, ex_layout, ex_sym, d_ex_taps, seg, d_out, num);
    }
    long long done = 0;
    last_kernel = "fir_direct";
    static const bool no_fork = getenv("SDR_B200_NOFORK") != nullptr;   // debugging aid: keep the tail on the main stream
    const bool can_fork = ctx->override_st == nullptr && !no_fork;
    bool fork_recorded = false;
    {
        // tuned kernel over the part of the FIRST segment it can take; the rest (ragged tail, straddling windows)
        // is finished by the generic kernel in the same tap order.  The tail only depends on the INPUT, so it runs
        // on the side stream concurrently with the tuned kernel (fork before, join after).
        // (no event when the tuned kernel will take everything: an event between two passes would also keep the next
        // pass from being scheduled early, see the programmatic dependent launch in kernels_fast.cu)
        const bool covers = D > 2 && dec_fast_will_cover(cplx, T, D, seg, num);
        if (can_fork && !covers) { SDR_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream)); fork_recorded = true; }
        const char *name = nullptr;
        if (D <= 2) SDR_TRY(launch_fir_small_stride_fast(ctx, cplx, T, D, d_taps, seg.a, seg.na, d_out, num, &done, &name));
        else        SDR_TRY(launch_dec_fast(ctx, cplx, T, D, d_taps, seg, d_out, num, &done, &name, h_taps.empty() ? nullptr : h_taps.data()));
        if (done > 0) last_kernel = name;
    }
    if (done == 0 && seg.nb > 0) {
        // Two segments and nothing for the tuned kernel in the first: a stage's short carried tail in front of vectors it
        // reads in place.  Only the few windows that START in the tail straddle the boundary -- the generic kernel
        // computes those (plus at most a few more, up to a 16-byte aligned start in the second segment); everything
        // after them is a one-segment problem.
        const long long m_a = (seg.na + D - 1) / D;
        long long m1 = -1;
        for (long long m = m_a; m < m_a + 16; m++)
            if (((((uintptr_t)seg.b) + (size_t)(m * D - seg.na) * eb) & 15) == 0) { m1 = m; break; }
        if (m1 >= 0 && m1 < num && m1 * D - seg.na < seg.nb) {
            if (m1 > 0) SDR_TRY(launch_fir_generic(ctx, cplx, T, D, d_taps, seg, d_out, m1));
            const long long off = m1 * D - seg.na;
            Seg2 rest = {(const char *)seg.b + (size_t)off * eb, seg.nb - off, nullptr, 0};
            return run(rest, 0, (char *)d_out + (size_t)m1 * eb, num - m1, cross_order);
        }
    }
    if (done < num) {
        Seg2 rest = seg;
        long long skip = done * D;
        if (skip >= rest.na) { rest.a = (const char *)rest.b + (skip - rest.na) * eb; rest.na = rest.nb - (skip - rest.na);
                               rest.b = nullptr; rest.nb = 0; if (rest.na < 0) rest.na = 0; }
        else                 { rest.a = (const char *)rest.a + skip * eb; rest.na -= skip; }
        const bool fork = can_fork && done > 0 && fork_recorded;
        if (fork) { SDR_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0)); ctx->override_st = ctx->side; }
        int rc = launch_fir_generic(ctx, cplx, T, D, d_taps, rest, (char *)d_out + done * eb, num - done);
        if (fork) {
            ctx->override_st = nullptr;
            if (rc == SDR_OK) { SDR_CUDA(cudaEventRecord(ctx->ev_join, ctx->side)); SDR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0)); }
        }
        SDR_TRY(rc);
    }
    return SDR_OK;
}

int FirRec::run_tuned(const void *d_in, long long n_in, long long first, void *d_out, long long num, long long *done) {
    *done = 0;
    if (num <= 0 || D <= 2 || arith != SDR_ARITH_FAST) return SDR_OK;
    const char *name = nullptr;
    // second segment with zero elements: keeps the kernel in its interior-only mode unless the whole request is resident
    Seg2 seg = {(const char *)d_in + first * elem_bytes(cplx), n_in - first, nullptr, 0};
    if ((num - 1) * (long long)D + T > seg.na) {   // not everything is resident: interior sub-tiles only
        long long fit = seg.na >= T ? (seg.na - T) / D + 1 : 0;
        if (fit < num) num = fit;
    }
    SDR_TRY(launch_dec_fast(ctx, cplx, T, D, d_taps, seg, d_out, num, done, &name, h_taps.empty() ? nullptr : h_taps.data()));
    if (*done > 0) last_kernel = name;
    return SDR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// ResRec: prepareCoeffs (FilterInternal.hs:297-319) restated as closed-form phase bookkeeping
// ---------------------------------------------------------------------------------------------------------------
int ResRec::create(Ctx *c, bool is_complex, int interpolation, int decimation, const float *coeffs, int n,
                   int size_multiple) {
    if (!c || !coeffs || n <= 0 || interpolation <= 0 || decimation <= 0 || size_multiple <= 0)
        return set_error(SDR_EINVAL, "resampler constructor: bad argument");
    ctx = c; cplx = is_complex; L = interpolation; M = decimation; n_taps = n; arith = c->arith;
    T = round_up(n, L * size_multiple);   // Filter.hs:422
    SDR_TRY(c->bind());
    // walk the phases in the order the reference visits them: offset' = L-1-((M-offset-1) mod L), starting at 0
    inc.clear(); prefix.clear(); offset_of_group.clear();
    int offset = 0;
    do {
        int d = M - offset - 1;
        // Haskell divMod floors; d can be negative when L > M
        int qd = (d >= 0) ? d / L : -((-d + L - 1) / L);
        int rd = d - qd * L;
        offset_of_group.push_back(offset);
        inc.push_back(qd + 1);
        offset = L - 1 - rd;
    } while (offset != 0 && (int)inc.size() <= L);
    ng = (int)inc.size();
    group_len = 0;
    for (int g = 0; g < ng; g++) {
        int len = (n - offset_of_group[g] + L - 1) / L;
        if (len < 0) len = 0;
        if (len > group_len) group_len = len;
    }
    row_stride = round_up(group_len, size_multiple);
    std::vector<float> table((size_t)ng * row_stride, 0.0f);
    sum_inc = 0;
    for (int g = 0; g < ng; g++) {
        prefix.push_back(sum_inc);
        sum_inc += inc[g];
        for (int l = 0, j = offset_of_group[g]; j < n; l++, j += L) table[(size_t)g * row_stride + l] = coeffs[j];
    }
    SDR_CUDA(cudaMalloc(&d_table, sizeof(float) * table.size()));
    SDR_CUDA(cudaMalloc(&d_prefix, sizeof(int) * ng));
    h_plain.assign(coeffs, coeffs + n);
    SDR_CUDA(cudaMalloc(&d_plain, sizeof(float) * n));
    SDR_CUDA(cudaMemcpyAsync(d_plain, coeffs, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    SDR_CUDA(cudaMemcpyAsync(d_table, table.data(), sizeof(float) * table.size(), cudaMemcpyHostToDevice, c->stream));
    SDR_CUDA(cudaMemcpyAsync(d_prefix, prefix.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, c->stream));
    SDR_CUDA(cudaStreamSynchronize(c->stream));
    return SDR_OK;
}

void ResRec::destroy() {
    if (ctx) { ctx->bind(); if (d_table) cudaFree(d_table); if (d_prefix) cudaFree(d_prefix); if (d_plain) cudaFree(d_plain); }
    d_table = nullptr; d_prefix = nullptr; d_plain = nullptr;
}

int ResRec::group_of_offset(int offset) const {
    for (int g = 0; g < ng; g++) if (offset_of_group[g] == offset) return g;
    return -1;
}

int ResRec::run(Seg2 seg, long long first, int g0, void *d_out, long long num, bool cross_order) {
    if (num <= 0) return SDR_OK;
    const size_t eb = elem_bytes(cplx);
    if (first >= seg.na) { seg.a = (const char *)seg.b + (first - seg.na) * eb; seg.na = seg.nb - (first - seg.na);
                           seg.b = nullptr; seg.nb = 0; if (seg.na < 0) seg.na = 0; }
    else                 { seg.a = (const char *)seg.a + first * eb; seg.na -= first; }
    if (arith == SDR_ARITH_EXACT) {
        // resampleAVXRR (avx_dotprod_R, W=8) / resampleAVXRC (avx_dotprod_C "2" form); cross kernels: left-to-right
        int W = cross_order ? 1 : 8, layout = cplx ? 2 : 0;
        int nt = cross_order ? group_len : row_stride;
        return launch_resample_exact(ctx, cplx, nt, row_stride, W, layout, g0, ng, d_prefix, sum_inc, d_table, seg, d_out, num);
    }
    last_kernel = "fir_tile";
    // tuned kernel for real streams: it starts on a cycle boundary (phase 0) with a 16-byte aligned window, so a short
    // generic prefix brings the stream there, the tuned kernel takes the bulk, the generic kernel the ragged end
    static const bool no_tuned = getenv("SDR_B200_NOTUNED") != nullptr;   // debugging aid
    if (ng == L && num >= 4096 && !no_tuned) {
        auto span = [&](long long count) { long long gi = (long long)g0 + count;
                                           return (gi / ng) * sum_inc + prefix[gi % ng] - prefix[g0]; };
        // first output (on a cycle boundary) whose window starts on a 16-byte boundary; with a short first segment (a
        // stage's carried tail in front of vectors read in place) the start is looked for in the second one
        long long p0 = -1;
        const bool tail_first = seg.nb > 0 && seg.na < 4096;
        const long long p_limit = 4LL * ng + (tail_first ? (seg.na * L) / M + 2LL * ng : 0);
        const char *t_ptr = nullptr;
        long long t_avail = 0;
        for (long long p = 0; p <= p_limit; p++) {
            if ((g0 + p) % ng != 0) continue;
            const long long sp = span(p);
            const char *ptr;
            long long avail;
            if (sp < seg.na) { if (tail_first) continue; ptr = (const char *)seg.a + (size_t)sp * eb; avail = seg.na - sp; }
            else if (seg.nb > 0 && sp - seg.na < seg.nb) { ptr = (const char *)seg.b + (size_t)(sp - seg.na) * eb; avail = seg.nb - (sp - seg.na); }
            else break;
            if ((((uintptr_t)ptr) & 15) == 0) { p0 = p; t_ptr = ptr; t_avail = avail; break; }
        }
        if (p0 >= 0 && p0 < num) {
            long long done = 0;
            const char *name = nullptr;
            const bool can_fork = ctx->override_st == nullptr;
            if (can_fork) SDR_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
            SDR_TRY(launch_res_fast(ctx, cplx, L, M, n_taps, d_plain, t_ptr, t_avail, (char *)d_out + (size_t)p0 * eb, num - p0, &done, &name));
            if (done > 0) {
                last_kernel = name;
                const bool fork = can_fork;
                if (fork) { SDR_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0)); ctx->override_st = ctx->side; }
                int rc = SDR_OK;
                if (p0 > 0) rc = launch_resample_groups(ctx, cplx, group_len, row_stride, g0, ng, d_prefix, sum_inc, d_table, seg, d_out, p0);
                long long k2 = p0 + done;   // first output of the ragged end; (g0 + k2) % ng == 0 again
                if (rc == SDR_OK && k2 < num) {
                    Seg2 rest = seg;
                    long long skip = span(k2);
                    if (skip >= rest.na) { rest.a = (const char *)rest.b + (skip - rest.na) * eb; rest.na = rest.nb - (skip - rest.na);
                                           rest.b = nullptr; rest.nb = 0; if (rest.na < 0) rest.na = 0; }
                    else                 { rest.a = (const char *)rest.a + skip * eb; rest.na -= skip; }
                    rc = launch_resample_groups(ctx, cplx, group_len, row_stride, (int)((g0 + k2) % ng), ng, d_prefix, sum_inc, d_table,
                                                rest, (char *)d_out + k2 * eb, num - k2);
                }
                if (fork) {
                    ctx->override_st = nullptr;
                    if (rc == SDR_OK) { SDR_CUDA(cudaEventRecord(ctx->ev_join, ctx->side)); SDR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0)); }
                }
                return rc;
            }
        }
    }
    return launch_resample_groups(ctx, cplx, group_len, row_stride, g0, ng, d_prefix, sum_inc, d_table, seg, d_out, num);
}

// ---------------------------------------------------------------------------------------------------------------
// one-shot helpers
// ---------------------------------------------------------------------------------------------------------------
enum { K_PLAIN = 0, K_DUP = 1, K_SYM = 2 };

static int default_arith() {
    const char *e = getenv("SDR_B200_ARITH");
    return (e && (!strcmp(e, "exact") || !strcmp(e, "EXACT") || !strcmp(e, "1"))) ? SDR_ARITH_EXACT : SDR_ARITH_FAST;
}

// taps + input staged in one H2D region: [taps (padded to 256 B)] [input]
static int oneshot_fir(const char *who, int kind, bool cplx, int num, int factor, int numCoeffs, const float *coeffs,
                       const float *in, float *out, int force_exact_variant /* -1: ctx default */) {
    if (num < 0 || factor <= 0 || numCoeffs <= 0 || !coeffs || (!in && num) || (!out && num))
        return set_error(SDR_EINVAL, "%s: bad argument (num %d, factor %d, numCoeffs %d)", who, num, factor, numCoeffs);
    if (kind == K_DUP && (numCoeffs & 1))
        return set_error(SDR_EINVAL, "%s: duplicated-coefficient form needs an even numCoeffs (got %d)", who, numCoeffs);
    if (num == 0) return SDR_OK;
    int st; Ctx *c = default_ctx(&st);
    if (st != SDR_OK) return st;
    std::vector<float> full;
    int T;
    if (kind == K_DUP)      { T = numCoeffs / 2; full.resize(T); for (int k = 0; k < T; k++) full[k] = coeffs[2 * k]; }
    else if (kind == K_SYM) { T = 2 * numCoeffs; full.assign(coeffs, coeffs + numCoeffs);
                              for (int k = numCoeffs - 1; k >= 0; k--) full.push_back(coeffs[k]); }
    else                    { T = numCoeffs; full.assign(coeffs, coeffs + numCoeffs); }
    const size_t eb = elem_bytes(cplx);
    long long n_in = (long long)(num - 1) * factor + T;
    full.resize(round_up(T, 128) + 128, 0.0f);   // zero-padded: the tuned kernels read their whole tap capacity
    size_t taps_bytes = ((sizeof(float) * full.size() + 255) / 256) * 256;
    size_t in_bytes = (size_t)n_in * eb, out_bytes = (size_t)num * eb;
    SDR_TRY(c->ensure_stage(taps_bytes + in_bytes, out_bytes));
    char *d_base = (char *)c->d_stage_in;
    SDR_CUDA(cudaMemcpyAsync(d_base, full.data(), sizeof(float) * full.size(), cudaMemcpyHostToDevice, c->stream));
    SDR_CUDA(cudaMemcpyAsync(d_base + taps_bytes, in, in_bytes, cudaMemcpyHostToDevice, c->stream));
    FirRec r;
    r.ctx = c; r.cplx = cplx; r.D = factor; r.T = T; r.d_taps = (float *)d_base; r.d_ex_taps = r.d_taps;
    r.h_taps.assign(full.begin(), full.begin() + T);
    r.arith = (force_exact_variant >= 0) ? SDR_ARITH_EXACT : default_arith();
    // EXACT: the AVX member of the family this entry point stands in for
    r.ex_W = 8; r.ex_sym = (kind == K_SYM); r.ex_T = (kind == K_SYM) ? numCoeffs : T;
    r.ex_layout = !cplx ? 0 : (kind == K_DUP ? 1 : 2);
    if (force_exact_variant >= 0) {
        switch (force_exact_variant) {
        case SDR_V_SCALAR: r.ex_W = 1; r.ex_layout = cplx ? 2 : 0; break;
        case SDR_V_SSE:    r.ex_W = 4; r.ex_layout = cplx ? 1 : 0; break;
        case SDR_V_AVX:    r.ex_W = 8; r.ex_layout = cplx ? 1 : 0; break;
        case SDR_V_SSE2:   r.ex_W = 4; r.ex_layout = 2; break;
        case SDR_V_AVX2:   r.ex_W = 8; r.ex_layout = 2; break;
        case SDR_V_SSESYM: r.ex_W = 4; r.ex_layout = cplx ? 2 : 0; break;
        case SDR_V_AVXSYM: r.ex_W = 8; r.ex_layout = cplx ? 2 : 0; break;
        default: return set_error(SDR_EINVAL, "%s: unknown variant %d", who, force_exact_variant);
        }
    }
    Seg2 seg = {d_base + taps_bytes, n_in, nullptr, 0};
    SDR_TRY(r.run(seg, 0, c->d_stage_out, num, false));
    SDR_CUDA(cudaMemcpyAsync(out, c->d_stage_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    SDR_CUDA(cudaStreamSynchronize(c->stream));
    return SDR_OK;
}

static int oneshot_resample(const char *who, bool cplx, int buf_size, int num_coeffs, int starting_group, int num_groups,
                            const int *increments, const float *const *rows, const float *flat, int flat_stride,
                            const float *in_buf, float *out_buf, int *next_group, int exact_variant) {
    if (buf_size < 0 || num_coeffs <= 0 || num_groups <= 0 || starting_group < 0 || starting_group >= num_groups ||
        !increments || (!rows && !flat) || (!in_buf && buf_size) || (!out_buf && buf_size))
        return set_error(SDR_EINVAL, "%s: bad argument", who);
    if (next_group) *next_group = (int)(((long long)starting_group + buf_size) % num_groups);
    if (buf_size == 0) return SDR_OK;
    int st; Ctx *c = default_ctx(&st);
    if (st != SDR_OK) return st;
    // input span: walk the increments once around, then closed form
    std::vector<int> prefix(num_groups);
    long long sum_inc = 0;
    for (int g = 0; g < num_groups; g++) { prefix[g] = (int)sum_inc; sum_inc += increments[g]; }
    long long gi_last = (long long)starting_group + buf_size - 1;
    long long start_last = (gi_last / num_groups) * sum_inc + prefix[gi_last % num_groups] - prefix[starting_group];
    long long n_in = start_last + num_coeffs;
    const size_t eb = elem_bytes(cplx);
    std::vector<float> table((size_t)num_groups * num_coeffs);
    for (int g = 0; g < num_groups; g++)
        memcpy(&table[(size_t)g * num_coeffs], rows ? rows[g] : flat + (size_t)g * flat_stride, sizeof(float) * num_coeffs);
    size_t table_bytes = ((table.size() * 4 + 255) / 256) * 256, prefix_bytes = ((num_groups * 4 + 255) / 256) * 256;
    size_t in_bytes = (size_t)n_in * eb, out_bytes = (size_t)buf_size * eb;
    SDR_TRY(c->ensure_stage(table_bytes + prefix_bytes + in_bytes, out_bytes));
    char *d_base = (char *)c->d_stage_in;
    SDR_CUDA(cudaMemcpyAsync(d_base, table.data(), table.size() * 4, cudaMemcpyHostToDevice, c->stream));
    SDR_CUDA(cudaMemcpyAsync(d_base + table_bytes, prefix.data(), num_groups * 4, cudaMemcpyHostToDevice, c->stream));
    SDR_CUDA(cudaMemcpyAsync(d_base + table_bytes + prefix_bytes, in_buf, in_bytes, cudaMemcpyHostToDevice, c->stream));
    Seg2 seg = {d_base + table_bytes + prefix_bytes, n_in, nullptr, 0};
    int arith = (exact_variant >= 0) ? SDR_ARITH_EXACT : default_arith();
    if (arith == SDR_ARITH_EXACT) {
        int W = 8, layout = cplx ? 2 : 0;
        if (exact_variant == SDR_V_SCALAR) W = 1;
        else if (exact_variant == SDR_V_SSE || exact_variant == SDR_V_SSE2) W = 4;
        else if (exact_variant >= 0 && exact_variant != SDR_V_AVX && exact_variant != SDR_V_AVX2)
            return set_error(SDR_EINVAL, "%s: variant %d has no resampler", who, exact_variant);
        SDR_TRY(launch_resample_exact(c, cplx, num_coeffs, num_coeffs, W, layout, starting_group, num_groups,
                                      (const int *)(d_base + table_bytes), (int)sum_inc, (const float *)d_base, seg,
                                      c->d_stage_out, buf_size));
    } else {
        SDR_TRY(launch_resample_groups(c, cplx, num_coeffs, num_coeffs, starting_group, num_groups,
                                       (const int *)(d_base + table_bytes), (int)sum_inc, (const float *)d_base, seg,
                                       c->d_stage_out, buf_size));
    }
    SDR_CUDA(cudaMemcpyAsync(out_buf, c->d_stage_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    SDR_CUDA(cudaStreamSynchronize(c->stream));
    return SDR_OK;
}

template <typename F>
static int oneshot_map(const char *who, long long n, size_t in_elem, size_t out_elem, const void *in, void *out, F body) {
    if (n < 0 || (n && (!in || !out))) return set_error(SDR_EINVAL, "%s: bad argument", who);
    if (n == 0) return SDR_OK;
    int st; Ctx *c = default_ctx(&st);
    if (st != SDR_OK) return st;
    Staged s = {c, SDR_HOST, in, out, (size_t)n * in_elem, (size_t)n * out_elem};
    SDR_TRY(s.begin());
    SDR_TRY(body(c, s.d_in, s.d_out));
    return s.end();
}

}  // namespace sdr

using namespace sdr;

extern "C" {

// ---- layer 1 ---------------------------------------------------------------------------------------------------
int filterCudaRR(int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    return oneshot_fir("filterCudaRR", K_PLAIN, false, num, 1, numCoeffs, coeffs, inBuf, outBuf, -1);
}
int filterCudaSymmetricRR(int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    return oneshot_fir("filterCudaSymmetricRR", K_SYM, false, num, 1, numCoeffs, coeffs, inBuf, outBuf, -1);
}
int filterCudaRC(int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    return oneshot_fir("filterCudaRC", K_PLAIN, true, num, 1, numCoeffs, coeffs, inBuf, outBuf, -1);
}
int filterCudaRCDup(int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    return oneshot_fir("filterCudaRCDup", K_DUP, true, num, 1, numCoeffs, coeffs, inBuf, outBuf, -1);
}
int filterCudaSymmetricRC(int num, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    return oneshot_fir("filterCudaSymmetricRC", K_SYM, true, num, 1, numCoeffs, coeffs, inBuf, outBuf, -1);
}
int decimateCudaRR(int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    return oneshot_fir("decimateCudaRR", K_PLAIN, false, num, factor, numCoeffs, coeffs, inBuf, outBuf, -1);
}
int decimateCudaSymmetricRR(int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    return oneshot_fir("decimateCudaSymmetricRR", K_SYM, false, num, factor, numCoeffs, coeffs, inBuf, outBuf, -1);
}
int decimateCudaRC(int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    return oneshot_fir("decimateCudaRC", K_PLAIN, true, num, factor, numCoeffs, coeffs, inBuf, outBuf, -1);
}
int decimateCudaRCDup(int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    return oneshot_fir("decimateCudaRCDup", K_DUP, true, num, factor, numCoeffs, coeffs, inBuf, outBuf, -1);
}
int decimateCudaSymmetricRC(int num, int factor, int numCoeffs, const float *coeffs, const float *inBuf, float *outBuf) {
    return oneshot_fir("decimateCudaSymmetricRC", K_SYM, true, num, factor, numCoeffs, coeffs, inBuf, outBuf, -1);
}

int sdr_exact_decimate(int variant, int is_complex, int num, int factor, int numCoeffs, const float *coeffs,
                       const float *inBuf, float *outBuf) {
    int kind = K_PLAIN;
    if (variant == SDR_V_SSESYM || variant == SDR_V_AVXSYM) kind = K_SYM;
    else if (is_complex && (variant == SDR_V_SSE || variant == SDR_V_AVX)) kind = K_DUP;
    if (!is_complex && (variant == SDR_V_SSE2 || variant == SDR_V_AVX2))
        return set_error(SDR_EINVAL, "sdr_exact_decimate: variant %d exists for complex data only", variant);
    if (variant < 0 || variant > SDR_V_AVXSYM) return set_error(SDR_EINVAL, "sdr_exact_decimate: unknown variant %d", variant);
    return oneshot_fir("sdr_exact_decimate", kind, is_complex != 0, num, factor, numCoeffs, coeffs, inBuf, outBuf, variant);
}

// resample2RR / resampleSSERR / resampleAVXRR and the RC forms (resample.c:34-142): the reference's exact signature --
// 8 arguments, the RETURN VALUE is the next group (what FilterInternal.mkResampler binds, :335-342, 358-362).  A
// failure returns -(status) < 0 with the message in sdr_last_error(); the reference's C cannot fail (void of errors).
int resampleCudaRR(int buf_size, int num_coeffs, int starting_group, int num_groups, int *increments, float **coeffs,
                   float *in_buf, float *out_buf) {
    int next = 0;
    int rc = oneshot_resample("resampleCudaRR", false, buf_size, num_coeffs, starting_group, num_groups, increments, coeffs,
                              nullptr, 0, in_buf, out_buf, &next, -1);
    return rc == SDR_OK ? next : -rc;
}
int resampleCudaRC(int buf_size, int num_coeffs, int starting_group, int num_groups, int *increments, float **coeffs,
                   float *in_buf, float *out_buf) {
    int next = 0;
    int rc = oneshot_resample("resampleCudaRC", true, buf_size, num_coeffs, starting_group, num_groups, increments, coeffs,
                              nullptr, 0, in_buf, out_buf, &next, -1);
    return rc == SDR_OK ? next : -rc;
}
int sdr_exact_resample(int variant, int is_complex, int buf_size, int num_coeffs, int starting_group, int num_groups,
                       const int *increments, const float *table, int row_stride, const float *in_buf, float *out_buf,
                       int *next_group) {
    if (variant < 0 || variant > SDR_V_AVX2) return set_error(SDR_EINVAL, "sdr_exact_resample: variant %d has no resampler", variant);
    return oneshot_resample("sdr_exact_resample", is_complex != 0, buf_size, num_coeffs, starting_group, num_groups,
                            increments, nullptr, table, row_stride, in_buf, out_buf, next_group, variant);
}

// resampleRR (resample.c:16-32): taps coeffs[filter_offset + l*interpolation], phase recurrence per output
int resampleCudaLegacyRR(int buf_size, int coeff_size, int interpolation, int decimation, int filter_offset,
                         const float *coeffs, const float *in_buf, float *out_buf) {
    if (buf_size < 0 || coeff_size <= 0 || interpolation <= 0 || decimation <= 0 || filter_offset < 0 ||
        filter_offset >= interpolation || !coeffs)
        return set_error(SDR_EINVAL, "resampleCudaLegacyRR: bad argument");
    if (buf_size == 0) return SDR_OK;
    int st; Ctx *c = default_ctx(&st);
    if (st != SDR_OK) return st;
    ResRec r;
    int saved = c->arith; c->arith = default_arith();
    int rc = r.create(c, false, interpolation, decimation, coeffs, coeff_size, 1);
    c->arith = saved;
    if (rc != SDR_OK) return rc;
    int g0 = r.group_of_offset(filter_offset);
    if (g0 < 0) { r.destroy(); return set_error(SDR_EINVAL, "resampleCudaLegacyRR: filter_offset %d is not a reachable phase", filter_offset); }
    // exactly the samples the reference loop touches: max over the last cycle of (window start + that phase's taps)
    long long n_in = 0;
    for (long long i = buf_size > r.ng ? buf_size - r.ng : 0; i < buf_size; i++) {
        long long gi = (long long)g0 + i;
        int g = (int)(gi % r.ng);
        long long len = (coeff_size - r.offset_of_group[g] + interpolation - 1) / interpolation;
        long long end = (gi / r.ng) * r.sum_inc + r.prefix[g] - r.prefix[g0] + (len > 0 ? len : 0);
        if (end > n_in) n_in = end;
    }
    Staged s = {c, SDR_HOST, in_buf, out_buf, (size_t)n_in * 4, (size_t)buf_size * 4};
    rc = s.begin();
    if (rc == SDR_OK) { Seg2 seg = {s.d_in, n_in, nullptr, 0}; rc = r.run(seg, 0, g0, s.d_out, buf_size, true); }
    if (rc == SDR_OK) rc = s.end();
    r.destroy();
    return rc;
}

int convertCuda(int num, const uint8_t *in, float *out) {
    return oneshot_map("convertCuda", num, 1, 4, in, out, [&](Ctx *c, const void *di, void *dout) {
        return launch_convert_u8(c, (const uint8_t *)di, (float *)dout, num); });
}
int convertCudaBladeRF(int num, const int16_t *in, float *out) {
    return oneshot_map("convertCudaBladeRF", num, 2, 4, in, out, [&](Ctx *c, const void *di, void *dout) {
        return launch_convert_i16(c, (const int16_t *)di, (float *)dout, num); });
}
int convertCudaBladeRFTransmit(int num, const float *in, int16_t *out) {
    return oneshot_map("convertCudaBladeRFTransmit", num, 4, 2, in, out, [&](Ctx *c, const void *di, void *dout) {
        return launch_convert_tx(c, (const float *)di, (int16_t *)dout, num); });
}
int scaleCuda(int num, float factor, const float *in_buf, float *out_buf) {
    return oneshot_map("scaleCuda", num, 4, 4, in_buf, out_buf, [&](Ctx *c, const void *di, void *dout) {
        return launch_scale(c, factor, (const float *)di, (float *)dout, num); });
}
int fmDemodCuda(int num, float lastRe, float lastIm, const float *in, float *out) {
    return oneshot_map("fmDemodCuda", num, 8, 4, in, out, [&](Ctx *c, const void *di, void *dout) {
        return launch_fm_demod(c, lastRe, lastIm, (const float *)di, (float *)dout, num); });
}
int dcBlockerCuda(int num, float lastSample, float lastOutput, float *finalSample, float *finalOutput,
                  const float *inBuf, float *outBuf) {
    if (num < 0 || (num && (!inBuf || !outBuf))) return set_error(SDR_EINVAL, "dcBlockerCuda: bad argument");
    if (finalSample) *finalSample = lastSample;
    if (finalOutput) *finalOutput = lastOutput;
    if (num == 0) return SDR_OK;
    int st; Ctx *c = default_ctx(&st);
    if (st != SDR_OK) return st;
    Staged s = {c, SDR_HOST, inBuf, outBuf, (size_t)num * 4, (size_t)num * 4 + 8};
    // the two final-state floats ride behind the output block
    SDR_TRY(c->bind());
    SDR_TRY(c->ensure_stage(s.in_bytes, s.out_bytes));
    SDR_CUDA(cudaMemcpyAsync(c->d_stage_in, inBuf, s.in_bytes, cudaMemcpyHostToDevice, c->stream));
    float *d_out = (float *)c->d_stage_out;
    SDR_TRY(launch_dc_blocker(c, lastSample, lastOutput, (const float *)c->d_stage_in, d_out, num, d_out + num));
    float fin[2];
    SDR_CUDA(cudaMemcpyAsync(outBuf, d_out, (size_t)num * 4, cudaMemcpyDeviceToHost, c->stream));
    SDR_CUDA(cudaMemcpyAsync(fin, d_out + num, 8, cudaMemcpyDeviceToHost, c->stream));
    SDR_CUDA(cudaStreamSynchronize(c->stream));
    if (finalSample) *finalSample = fin[0];
    if (finalOutput) *finalOutput = fin[1];
    return SDR_OK;
}

// ---- layer 2: Filter ---------------------------------------------------------------------------------------------
static Ctx *as_ctx(sdr_ctx_t *c) { return reinterpret_cast<Ctx *>(c); }

int sdr_filter_create(sdr_ctx_t *ctx, int is_complex, const float *coeffs, int num_coeffs, int size_multiple, sdr_filter_t **f) {
    if (!f) return set_error(SDR_EINVAL, "sdr_filter_create: null out pointer");
    *f = nullptr;
    sdr_filter *h = new sdr_filter();
    int rc = h->r.create(as_ctx(ctx), is_complex != 0, 1, coeffs, num_coeffs, size_multiple, false);
    if (rc != SDR_OK) { delete h; return rc; }
    *f = h;
    return SDR_OK;
}
int sdr_filter_create_sym(sdr_ctx_t *ctx, int is_complex, const float *half_coeffs, int half_len, sdr_filter_t **f) {
    if (!f) return set_error(SDR_EINVAL, "sdr_filter_create_sym: null out pointer");
    *f = nullptr;
    sdr_filter *h = new sdr_filter();
    int rc = h->r.create(as_ctx(ctx), is_complex != 0, 1, half_coeffs, half_len, 1, true);
    if (rc != SDR_OK) { delete h; return rc; }
    *f = h;
    return SDR_OK;
}
int sdr_filter_destroy(sdr_filter_t *f) { if (f) { f->r.destroy(); delete f; } return SDR_OK; }
int sdr_filter_num_coeffs(const sdr_filter_t *f) { return f ? f->r.T : -1; }
const char *sdr_filter_last_kernel(const sdr_filter_t *f) { return f ? f->r.last_kernel : "none"; }

static int rec_one(FirRec &r, const char *who, int count, const void *in, void *out, int mem) {
    if (count < 0 || (count && (!in || !out)) || (mem != SDR_HOST && mem != SDR_DEVICE))
        return set_error(SDR_EINVAL, "%s: bad argument", who);
    if (count == 0) return SDR_OK;
    const size_t eb = elem_bytes(r.cplx);
    long long n_in = (long long)(count - 1) * r.D + r.T;
    Staged s = {r.ctx, mem, in, out, (size_t)n_in * eb, (size_t)count * eb};
    SDR_TRY(s.begin());
    Seg2 seg = {s.d_in, n_in, nullptr, 0};
    SDR_TRY(r.run(seg, 0, s.d_out, count, false));
    return s.end();
}

static int rec_cross(FirRec &r, const char *who, int count, const void *last, int n_last, const void *next, int n_next,
                     void *out, int mem) {
    if (count < 0 || n_last < 0 || n_next < 0 || (count && !out) || (n_last && !last) || (n_next && !next) ||
        (mem != SDR_HOST && mem != SDR_DEVICE))
        return set_error(SDR_EINVAL, "%s: bad argument", who);
    if (count == 0) return SDR_OK;
    const size_t eb = elem_bytes(r.cplx);
    // only the part of `next` the windows can reach is touched (and staged)
    long long reach = (long long)(count - 1) * r.D + r.T - n_last;
    long long use_next = reach < 0 ? 0 : (reach < n_next ? reach : n_next);
    SDR_TRY(r.ctx->bind());
    if (mem == SDR_DEVICE) {
        Seg2 seg = {last, n_last, next, use_next};
        return r.run(seg, 0, out, count, true);
    }
    size_t in_bytes = (size_t)(n_last + use_next) * eb, out_bytes = (size_t)count * eb;
    SDR_TRY(r.ctx->ensure_stage(in_bytes, out_bytes));
    char *d = (char *)r.ctx->d_stage_in;
    if (n_last) SDR_CUDA(cudaMemcpyAsync(d, last, (size_t)n_last * eb, cudaMemcpyHostToDevice, r.ctx->stream));
    if (use_next) SDR_CUDA(cudaMemcpyAsync(d + (size_t)n_last * eb, next, (size_t)use_next * eb, cudaMemcpyHostToDevice, r.ctx->stream));
    Seg2 seg = {d, n_last + use_next, nullptr, 0};
    SDR_TRY(r.run(seg, 0, r.ctx->d_stage_out, count, true));
    SDR_CUDA(cudaMemcpyAsync(out, r.ctx->d_stage_out, out_bytes, cudaMemcpyDeviceToHost, r.ctx->stream));
    SDR_CUDA(cudaStreamSynchronize(r.ctx->stream));
    return SDR_OK;
}

int sdr_filter_one(sdr_filter_t *f, int count, const void *in, void *out, int mem) {
    if (!f) return set_error(SDR_EINVAL, "sdr_filter_one: null handle");
    return rec_one(f->r, "sdr_filter_one", count, in, out, mem);
}
int sdr_filter_cross(sdr_filter_t *f, int count, const void *last, int n_last, const void *next, int n_next, void *out, int mem) {
    if (!f) return set_error(SDR_EINVAL, "sdr_filter_cross: null handle");
    return rec_cross(f->r, "sdr_filter_cross", count, last, n_last, next, n_next, out, mem);
}

// ---- layer 2: Decimator ------------------------------------------------------------------------------------------
int sdr_decimator_create(sdr_ctx_t *ctx, int is_complex, int factor, const float *coeffs, int num_coeffs,
                         int size_multiple, sdr_decimator_t **d) {
    if (!d) return set_error(SDR_EINVAL, "sdr_decimator_create: null out pointer");
    *d = nullptr;
    sdr_decimator *h = new sdr_decimator();
    int rc = h->r.create(as_ctx(ctx), is_complex != 0, factor, coeffs, num_coeffs, size_multiple, false);
    if (rc != SDR_OK) { delete h; return rc; }
    *d = h;
    return SDR_OK;
}
int sdr_decimator_create_sym(sdr_ctx_t *ctx, int is_complex, int factor, const float *half_coeffs, int half_len,
                             sdr_decimator_t **d) {
    if (!d) return set_error(SDR_EINVAL, "sdr_decimator_create_sym: null out pointer");
    *d = nullptr;
    sdr_decimator *h = new sdr_decimator();
    int rc = h->r.create(as_ctx(ctx), is_complex != 0, factor, half_coeffs, half_len, 1, true);
    if (rc != SDR_OK) { delete h; return rc; }
    *d = h;
    return SDR_OK;
}
int sdr_decimator_destroy(sdr_decimator_t *d) { if (d) { d->r.destroy(); delete d; } return SDR_OK; }
int sdr_decimator_num_coeffs(const sdr_decimator_t *d) { return d ? d->r.T : -1; }
int sdr_decimator_factor(const sdr_decimator_t *d) { return d ? d->r.D : -1; }
const char *sdr_decimator_last_kernel(const sdr_decimator_t *d) { return d ? d->r.last_kernel : "none"; }

int sdr_decimate_one(sdr_decimator_t *d, int count, const void *in, void *out, int mem) {
    if (!d) return set_error(SDR_EINVAL, "sdr_decimate_one: null handle");
    return rec_one(d->r, "sdr_decimate_one", count, in, out, mem);
}
int sdr_decimate_cross(sdr_decimator_t *d, int count, const void *last, int n_last, const void *next, int n_next,
                       void *out, int mem) {
    if (!d) return set_error(SDR_EINVAL, "sdr_decimate_cross: null handle");
    return rec_cross(d->r, "sdr_decimate_cross", count, last, n_last, next, n_next, out, mem);
}
// whole device-resident stream in one call: y[m], m < num, over n_in resident samples (bench.py, multi-GPU interior)
int sdr_decimate_stream(sdr_decimator_t *d, const void *d_in, long long n_in, void *d_out, long long num) {
    if (!d || num < 0 || n_in < 0 || (num && (!d_in || !d_out))) return set_error(SDR_EINVAL, "sdr_decimate_stream: bad argument");
    if (num && (num - 1) * d->r.D + d->r.T > n_in)
        return set_error(SDR_EPRECOND, "sdr_decimate_stream: %lld outputs need %lld samples, %lld resident", num,
                         (num - 1) * d->r.D + d->r.T, n_in);
    SDR_TRY(d->r.ctx->bind());
    Seg2 seg = {d_in, n_in, nullptr, 0};
    return d->r.run(seg, 0, d_out, num, false);
}

int sdr_filter_stream(sdr_filter_t *f, const void *d_in, long long n_in, void *d_out, long long num) {
    if (!f || num < 0 || n_in < 0 || (num && (!d_in || !d_out))) return set_error(SDR_EINVAL, "sdr_filter_stream: bad argument");
    if (num && (num - 1) + f->r.T > n_in)
        return set_error(SDR_EPRECOND, "sdr_filter_stream: %lld outputs need %lld samples, %lld resident", num, (num - 1) + f->r.T, n_in);
    SDR_TRY(f->r.ctx->bind());
    Seg2 seg = {d_in, n_in, nullptr, 0};
    return f->r.run(seg, 0, d_out, num, false);
}

// ---- layer 2: Resampler ------------------------------------------------------------------------------------------
int sdr_resampler_create(sdr_ctx_t *ctx, int is_complex, int interpolation, int decimation, const float *coeffs,
                         int num_coeffs, int size_multiple, sdr_resampler_t **r) {
    if (!r) return set_error(SDR_EINVAL, "sdr_resampler_create: null out pointer");
    *r = nullptr;
    sdr_resampler *h = new sdr_resampler();
    int rc = h->r.create(as_ctx(ctx), is_complex != 0, interpolation, decimation, coeffs, num_coeffs, size_multiple);
    if (rc != SDR_OK) { delete h; return rc; }
    *r = h;
    return SDR_OK;
}
int sdr_resampler_destroy(sdr_resampler_t *r) { if (r) { r->r.destroy(); delete r; } return SDR_OK; }
int sdr_resampler_num_coeffs(const sdr_resampler_t *r) { return r ? r->r.T : -1; }
const char *sdr_resampler_last_kernel(const sdr_resampler_t *r) { return r ? r->r.last_kernel : "none"; }
int sdr_resampler_interpolation(const sdr_resampler_t *r) { return r ? r->r.L : -1; }
int sdr_resampler_decimation(const sdr_resampler_t *r) { return r ? r->r.M : -1; }

// input samples consumed up to (not including) output `count` when starting at group g0
static long long res_span(const ResRec &r, int g0, long long count) {
    long long gi = (long long)g0 + count;
    return (gi / r.ng) * r.sum_inc + r.prefix[gi % r.ng] - r.prefix[g0];
}

int sdr_resample_one(sdr_resampler_t *h, sdr_resampler_dat_t *dat, int count, const void *in, void *out, int mem,
                     int *end_offset) {
    if (!h || !dat || count < 0 || (count && (!in || !out)) || (mem != SDR_HOST && mem != SDR_DEVICE))
        return set_error(SDR_EINVAL, "sdr_resample_one: bad argument");
    ResRec &r = h->r;
    if (dat->group < 0 || dat->group >= r.ng) return set_error(SDR_EINVAL, "sdr_resample_one: group %d out of range", dat->group);
    int g0 = dat->group;
    if (count) {
        const size_t eb = elem_bytes(r.cplx);
        long long n_in = res_span(r, g0, count - 1) + r.group_len;
        Staged s = {r.ctx, mem, in, out, (size_t)n_in * eb, (size_t)count * eb};
        SDR_TRY(s.begin());
        Seg2 seg = {s.d_in, n_in, nullptr, 0};
        SDR_TRY(r.run(seg, 0, g0, s.d_out, count, false));
        SDR_TRY(s.end());
    }
    // func1 Filter.hs:423: the C call returns the next group; the offset is recomputed from it
    int group = (int)(((long long)g0 + count) % r.ng);
    int offset = r.L - 1 - (int)(((long long)r.L + (long long)group * r.M - 1) % r.L);
    dat->group = group; dat->offset = offset;
    if (end_offset) *end_offset = offset;
    return SDR_OK;
}

// whole device-resident stream in one call, starting at output 0 of the stream (group 0, offset 0)
int sdr_resample_stream(sdr_resampler_t *h, const void *d_in, long long n_in, void *d_out, long long num) {
    if (!h || num < 0 || n_in < 0 || (num && (!d_in || !d_out))) return set_error(SDR_EINVAL, "sdr_resample_stream: bad argument");
    ResRec &r = h->r;
    if (num && res_span(r, 0, num - 1) + r.group_len > n_in)
        return set_error(SDR_EPRECOND, "sdr_resample_stream: %lld outputs need %lld samples, %lld resident", num,
                         res_span(r, 0, num - 1) + r.group_len, n_in);
    SDR_TRY(r.ctx->bind());
    Seg2 seg = {d_in, n_in, nullptr, 0};
    return r.run(seg, 0, 0, d_out, num, false);
}

int sdr_resample_cross(sdr_resampler_t *h, sdr_resampler_dat_t *dat, int count, const void *last, int n_last,
                       const void *next, int n_next, void *out, int mem, int *end_offset) {
    if (!h || !dat || count < 0 || n_last < 0 || n_next < 0 || (count && !out) || (mem != SDR_HOST && mem != SDR_DEVICE))
        return set_error(SDR_EINVAL, "sdr_resample_cross: bad argument");
    ResRec &r = h->r;
    int g0 = r.group_of_offset(dat->offset);
    if (g0 < 0) return set_error(SDR_EINVAL, "sdr_resample_cross: offset %d is not a reachable phase", dat->offset);
    if (count) {
        const size_t eb = elem_bytes(r.cplx);
        long long reach = res_span(r, g0, count - 1) + r.group_len - n_last;
        long long use_next = reach < 0 ? 0 : (reach < n_next ? reach : n_next);
        SDR_TRY(r.ctx->bind());
        if (mem == SDR_DEVICE) {
            Seg2 seg = {last, n_last, next, use_next};
            SDR_TRY(r.run(seg, 0, g0, out, count, true));
        } else {
            size_t in_bytes = (size_t)(n_last + use_next) * eb, out_bytes = (size_t)count * eb;
            SDR_TRY(r.ctx->ensure_stage(in_bytes, out_bytes));
            char *d = (char *)r.ctx->d_stage_in;
            if (n_last) SDR_CUDA(cudaMemcpyAsync(d, last, (size_t)n_last * eb, cudaMemcpyHostToDevice, r.ctx->stream));
            if (use_next) SDR_CUDA(cudaMemcpyAsync(d + (size_t)n_last * eb, next, (size_t)use_next * eb, cudaMemcpyHostToDevice, r.ctx->stream));
            Seg2 seg = {d, n_last + use_next, nullptr, 0};
            SDR_TRY(r.run(seg, 0, g0, r.ctx->d_stage_out, count, true));
            SDR_CUDA(cudaMemcpyAsync(out, r.ctx->d_stage_out, out_bytes, cudaMemcpyDeviceToHost, r.ctx->stream));
            SDR_CUDA(cudaStreamSynchronize(r.ctx->stream));
        }
    }
    // Filter.hs:421,440: group' = (group + count) mod interpolation; offset' from the phase recurrence
    int g_end = (int)(((long long)g0 + count) % r.ng);
    dat->group = (int)(((long long)dat->group + count) % r.L);
    dat->offset = r.offset_of_group[g_end];
    if (end_offset) *end_offset = dat->offset;
    return SDR_OK;
}

}  // extern "C"
