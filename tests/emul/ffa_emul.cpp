// TEST INFRASTRUCTURE: runs the product's 2-parallel fast-FIR lane function (sdr_b200/csrc/fir_ffa.cuh, the function the
// CUDA kernel k_fir_r_ffa_ring is unrolled from) on the host over a whole stream, lane by lane, so its index arithmetic
// and its accuracy are checked against the oracle without a GPU.  Built by tests/test_fir_ffa_model.py.
#include "../../sdr_b200/csrc/fir_ffa.cuh"

#include <string.h>

using namespace sdr;

template <int T, int R> static long long run(const float *x, long long n, const float *taps, float *y) {
    constexpr int H = T / 2, WIN4 = (R + T - 1 + 3) / 4;
    float h0[H], h1[H], hs[H];
    for (int j = 0; j < H; j++) { h0[j] = taps[2 * j]; h1[j] = taps[2 * j + 1]; hs[j] = h0[j] + h1[j]; }
    long long done = 0;
    for (long long o = 0; o + WIN4 * 4 <= n; o += R) {   // one "lane" per R outputs, window read in place
        float4 w[WIN4];
        memcpy(w, x + o, sizeof(w));
        float out[R];
        fir_ffa_lane<T, R>(w, h0, h1, hs, out);
        memcpy(y + o, out, sizeof(out));
        done = o + R;
    }
    return done;
}

extern "C" long long emul_fir_ffa(int T, const float *x, long long n, const float *taps, float *y) {
    if (T == 64) return run<64, 20>(x, n, taps, y);
    if (T == 32) return run<32, 20>(x, n, taps, y);
    return -1;
}
