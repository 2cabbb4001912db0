// launch_ring_real: real-data instantiations of the padded-segment ring decimator with the taps in registers (dec_ring.cuh)
#include "dec_ring.cuh"

namespace sdr {

int launch_ring_real(Ctx *c, int T, int D, const float *d_taps, Seg2 seg, void *d_out, long long num, long long *done, const char **name) {
    *done = 0;
    if (T == 128 && D == 8) { *name = "dec_r_ring<128,8,16>"; return launch_ring<false, 128, 8, 16>(c, d_taps, seg, d_out, num, done); }
    if (T == 64 && D == 8) { *name = "dec_r_ring<64,8,16>"; return launch_ring<false, 64, 8, 16>(c, d_taps, seg, d_out, num, done); }
    if (T == 32 && D == 8) { *name = "dec_r_ring<32,8,16>"; return launch_ring<false, 32, 8, 16>(c, d_taps, seg, d_out, num, done); }
    if (T == 128 && D == 4) { *name = "dec_r_ring<128,4,32>"; return launch_ring<false, 128, 4, 32>(c, d_taps, seg, d_out, num, done); }
    if (T == 64 && D == 4) { *name = "dec_r_ring<64,4,32>"; return launch_ring<false, 64, 4, 32>(c, d_taps, seg, d_out, num, done); }
    if (T == 32 && D == 4) { *name = "dec_r_ring<32,4,32>"; return launch_ring<false, 32, 4, 32>(c, d_taps, seg, d_out, num, done); }
    if (T == 128 && D == 16) { *name = "dec_r_ring<128,16,8>"; return launch_ring<false, 128, 16, 8>(c, d_taps, seg, d_out, num, done); }
    if (T == 64 && D == 16) { *name = "dec_r_ring<64,16,8>"; return launch_ring<false, 64, 16, 8>(c, d_taps, seg, d_out, num, done); }
    if (T == 32 && D == 16) { *name = "dec_r_ring<32,16,8>"; return launch_ring<false, 32, 16, 8>(c, d_taps, seg, d_out, num, done); }
    return SDR_OK;
}

}  // namespace sdr
