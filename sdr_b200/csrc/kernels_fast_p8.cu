// one instantiation: complex data, 256 taps as launch parameters, decimate by 8 (dec_ring.cuh; dispatched from kernels_fast_p.cu).
// On its own because the compiler front end spends two minutes on its fully unrolled 2048-FFMA2 blocks.
#include "dec_ring.cuh"

namespace sdr {

int launch_ring_param_c256_d8(Ctx *c, const float *d_taps, const float *h_taps, Seg2 seg, void *d_out, long long num, long long *done) {
    return launch_ring<true, 256, 8, 8, true>(c, d_taps, seg, d_out, num, done, h_taps);
}

}  // namespace sdr
