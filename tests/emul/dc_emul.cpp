// TEST INFRASTRUCTURE: runs the product's dcBlocker speculation (sdr_b200/csrc/dc_spec.cuh -- the very functions the
// CUDA kernels k_dc_spec / k_dc_repair execute, compiled here for the host) one chunk after the other, so the chunk
// planning, warm-up, verification and repair logic is checked bit for bit against the oracle without a GPU.
// Built by tests/test_dc_speculation.py with g++ -O2 -ffp-contract=off; never linked into libsdr_b200.so.
#include "../../sdr_b200/csrc/dc_spec.cuh"

#include <vector>

using namespace sdr;

extern "C" int emul_dc_blocker(const float *in, float *out, long long n, float last_sample, float last_output, int ch,
                               int k1, int k2, int vec, int reverse_chunks, float *final2, unsigned long long *stats, int tile) {
    if (n <= 0 || (tile != 32 && tile != 64) || ch < tile || (ch % tile) || (k1 % tile) || (k2 % tile)) return 1;
    DcArgs A;
    A.in = in; A.out = out; A.n = n;
    A.last_sample = last_sample; A.last_output = last_output; A.state_in = nullptr;
    A.ch = ch; A.k1 = k1; A.k2 = k2;
    A.chunks = (n + ch - 1) / ch;
    std::vector<uint32_t> spec(A.chunks, 0xdeadbeefu), fin(A.chunks, 0xdeadbeefu), bits((A.chunks + 31) / 32, 0);
    A.spec = spec.data(); A.fin = fin.data(); A.fail_bits = bits.data();
    A.stats = stats; A.final2 = final2;
    // the lanes of the kernel run in no particular order
    for (long long i = 0; i < A.chunks; i++) {
        const long long c = reverse_chunks ? A.chunks - 1 - i : i;
        // `vec` selects the arithmetic flavour here (0 widen both, 1 widen the difference, 2 native, 3 native incl. the
        // cheap warm-up), `tile` the walk granularity: identical bits for all of them
        if (tile == 64) dc_chunk<DC_NATIVE_ALL, 64>(A, c);
        else if (vec == 0) dc_chunk<DC_WIDEN_BOTH, 32>(A, c);
        else if (vec == 1) dc_chunk<DC_WIDEN_DIFF, 32>(A, c);
        else if (vec == 2) dc_chunk<DC_NATIVE, 32>(A, c);
        else dc_chunk<DC_NATIVE_ALL, 32>(A, c);
    }
    bool any = false;
    for (long long c = 1; c < A.chunks; c++)
        if (dc_missed(A, c, A.fin[c - 1])) { bits[c / 32] |= 1u << (c % 32); any = true; }
    if (any) dc_repair(A);
    else {
        stats[0] += 1; stats[1] += (unsigned long long)A.chunks;
        if (final2) { final2[0] = in[n - 1]; final2[1] = dc_float(A.fin[A.chunks - 1]); }
    }
    return 0;
}

// dc_widen against the compiler's own float -> double conversion over bit patterns first, first + step, ... (step 1 = all 2^32)
extern "C" unsigned long long emul_dc_widen_mismatches(unsigned long long first, unsigned long long step) {
    unsigned long long bad = 0;
    for (unsigned long long i = first; i < (1ull << 32); i += step) {
        const float  f = dc_float((uint32_t)i);
        const double want = (double)f, got = dc_widen(f);
        if (memcmp(&want, &got, 8) != 0) bad++;
    }
    return bad;
}
