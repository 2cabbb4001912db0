// records.cuh -- device-side state behind the plugin records (reference Filter.hs:116-144) shared by the record,
// pipe and multi-GPU layers.
#pragma once
#include <vector>

#include "common.cuh"

namespace sdr {

// One FIR / decimator: taps resident on the device in the forms the kernels want.
struct FirRec {
    Ctx  *ctx = nullptr;
    bool  cplx = false;
    int   D = 1;            // decimation (1 for filters)
    int   T = 0;            // numCoeffsF / numCoeffsD: stored (possibly zero-padded) tap count
    int   arith = SDR_ARITH_FAST;
    float *d_taps = nullptr;       // T floats, full (symmetric halves expanded), zero padded
    std::vector<float> h_taps;     // the same T floats on the host (kernels that take their taps as launch parameters)
    // EXACT arithmetic: the variant the reference's fast* constructor would pick on an AVX host, with the taps in the
    // layout that C function receives
    int   ex_W = 8, ex_layout = 0, ex_sym = 0, ex_T = 0;
    float *d_ex_taps = nullptr;
    const char *last_kernel = "none";
    bool  symmetric = false;   // c[k] == c[T-1-k] bit for bit (lets kernels keep half the taps in registers)
    int create(Ctx *c, bool is_complex, int factor, const float *coeffs, int n, int size_multiple, bool sym_half);
    void destroy();
    // y[m] = sum_k c[k] x[first + m*D + k], x = seg.a ++ seg.b (zeros beyond), m < num.  `first` in elements.
    int run(Seg2 seg, long long first, void *d_out, long long num, bool cross_order);
    // tuned kernel only, over the leading outputs it can take from the single segment (*done of them); nothing else
    int run_tuned(const void *d_in, long long n_in, long long first, void *d_out, long long num, long long *done);
};

struct ResRec {
    Ctx  *ctx = nullptr;
    bool  cplx = false;
    int   L = 1, M = 1;
    int   T = 0;               // numCoeffsR (Filter.hs:422)
    int   n_taps = 0;          // unpadded tap count
    int   ng = 0;              // polyphase groups (L / gcd(L, M))
    int   group_len = 0;       // max taps per group (what the reference passes to C as num_coeffs)
    int   row_stride = 0;      // group_len rounded up to the size multiple (zero padded)
    int   sum_inc = 0;
    int   arith = SDR_ARITH_FAST;
    std::vector<int> inc, prefix, offset_of_group;
    float *d_table = nullptr;  // [ng][row_stride]
    int   *d_prefix = nullptr; // [ng]
    float *d_plain = nullptr;  // the n_taps plain coefficients (tuned kernels index them as c[f + l L])
    std::vector<float> h_plain; // the same on the host
    const char *last_kernel = "none";
    int create(Ctx *c, bool is_complex, int interpolation, int decimation, const float *coeffs, int n, int size_multiple);
    void destroy();
    int group_of_offset(int offset) const;
    // outputs i < num: group (g0 + i) % ng, window starting at element `first` of seg for i = 0
    int run(Seg2 seg, long long first, int g0, void *d_out, long long num, bool cross_order);
};

static inline size_t elem_bytes(bool cplx) { return cplx ? 8 : 4; }

// Stage a host buffer to the device / a device buffer back to the host around `body` when mem == SDR_HOST.
struct Staged {
    Ctx *c; int mem; const void *in; void *out; size_t in_bytes, out_bytes;
    const void *d_in = nullptr; void *d_out = nullptr;
    int begin();   // after this d_in / d_out are device pointers
    int end();     // copies the result back and synchronises when mem == SDR_HOST
};

}  // namespace sdr

struct sdr_filter    { sdr::FirRec r; };
struct sdr_decimator { sdr::FirRec r; };
struct sdr_resampler { sdr::ResRec r; };
