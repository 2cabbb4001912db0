// launch_ring_complex: complex-data instantiations of the padded-segment ring decimator with the taps in registers (dec_ring.cuh)
#include "dec_ring.cuh"

namespace sdr {

int launch_ring_complex(Ctx *c, int T, int D, const float *d_taps, Seg2 seg, void *d_out, long long num, long long *done, const char **name) {
    *done = 0;
    if (T == 128 && D == 8) { *name = "dec_c_ring<128,8,8>"; return launch_ring<true, 128, 8, 8>(c, d_taps, seg, d_out, num, done); }
    if (T == 64 && D == 8) { *name = "dec_c_ring<64,8,8>"; return launch_ring<true, 64, 8, 8>(c, d_taps, seg, d_out, num, done); }
    if (T == 32 && D == 8) { *name = "dec_c_ring<32,8,8>"; return launch_ring<true, 32, 8, 8>(c, d_taps, seg, d_out, num, done); }
    if (T == 128 && D == 4) { *name = "dec_c_ring<128,4,16>"; return launch_ring<true, 128, 4, 16>(c, d_taps, seg, d_out, num, done); }
    if (T == 64 && D == 4) { *name = "dec_c_ring<64,4,16>"; return launch_ring<true, 64, 4, 16>(c, d_taps, seg, d_out, num, done); }
    if (T == 32 && D == 4) { *name = "dec_c_ring<32,4,16>"; return launch_ring<true, 32, 4, 16>(c, d_taps, seg, d_out, num, done); }
    if (T == 128 && D == 16) { *name = "dec_c_ring<128,16,4>"; return launch_ring<true, 128, 16, 4>(c, d_taps, seg, d_out, num, done); }
    if (T == 64 && D == 16) { *name = "dec_c_ring<64,16,4>"; return launch_ring<true, 64, 16, 4>(c, d_taps, seg, d_out, num, done); }
    if (T == 32 && D == 16) { *name = "dec_c_ring<32,16,4>"; return launch_ring<true, 32, 16, 4>(c, d_taps, seg, d_out, num, done); }
    return SDR_OK;
}

}  // namespace sdr
