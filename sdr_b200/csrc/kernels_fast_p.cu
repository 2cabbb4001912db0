// launch-parameter instantiations of the padded-segment ring decimator (dec_ring.cuh): the taps travel as a kernel parameter
// and reach every FFMA / FFMA2 as a uniform-register operand -- 129..256 taps, and the 16-warp real-data decimators
#include "dec_ring.cuh"

namespace sdr {

int launch_ring_param_c256_d8(Ctx *c, const float *d_taps, const float *h_taps, Seg2 seg, void *d_out, long long num, long long *done);
int launch_ring_param_c256_d4(Ctx *c, const float *d_taps, const float *h_taps, Seg2 seg, void *d_out, long long num, long long *done);

int launch_ring_param(Ctx *c, bool cplx, int T, int D, const float *d_taps, const float *h_taps, Seg2 seg, void *d_out, long long num,
                      long long *done, const char **name) {
    *done = 0;
    // (the two instantiations the compiler front end needs two minutes each for have a translation unit of their own: kernels_fast_p8.cu, _p4.cu)
    if (T == 256 && cplx == true && D == 8) { *name = "dec_c_ring<256,8,8,param>"; return launch_ring_param_c256_d8(c, d_taps, h_taps, seg, d_out, num, done); }
    if (T == 256 && cplx == true && D == 4) { *name = "dec_c_ring<256,4,8,param>"; return launch_ring_param_c256_d4(c, d_taps, h_taps, seg, d_out, num, done); }
    if (T == 256 && cplx == true && D == 16) { *name = "dec_c_ring<256,16,4,param>"; return launch_ring<true, 256, 16, 4, true>(c, d_taps, seg, d_out, num, done, h_taps); }
    if (T == 256 && cplx == false && D == 8) { *name = "dec_r_ring<256,8,8,param>"; return launch_ring<false, 256, 8, 8, true>(c, d_taps, seg, d_out, num, done, h_taps); }
    if (T == 256 && cplx == false && D == 4) { *name = "dec_r_ring<256,4,8,param>"; return launch_ring<false, 256, 4, 8, true>(c, d_taps, seg, d_out, num, done, h_taps); }
    if (T == 256 && cplx == false && D == 16) { *name = "dec_r_ring<256,16,8,param>"; return launch_ring<false, 256, 16, 8, true>(c, d_taps, seg, d_out, num, done, h_taps); }
    if (T == 128 && !cplx && D == 8) { *name = "dec_r_ring<128,8,8,param,16w>"; return launch_ring<false, 128, 8, 8, true, 16>(c, d_taps, seg, d_out, num, done, h_taps); }
    if (T == 128 && !cplx && D == 4) { *name = "dec_r_ring<128,4,8,param,16w>"; return launch_ring<false, 128, 4, 8, true, 16>(c, d_taps, seg, d_out, num, done, h_taps); }
    if (T == 64 && !cplx && D == 4) { *name = "dec_r_ring<64,4,8,param,16w>"; return launch_ring<false, 64, 4, 8, true, 16>(c, d_taps, seg, d_out, num, done, h_taps); }
    if (T == 64 && !cplx && D == 8) { *name = "dec_r_ring<64,8,8,param,16w>"; return launch_ring<false, 64, 8, 8, true, 16>(c, d_taps, seg, d_out, num, done, h_taps); }
    if (T == 128 && cplx && D == 4) { *name = "dec_c_ring<128,4,8,param,16w>"; return launch_ring<true, 128, 4, 8, true, 16>(c, d_taps, seg, d_out, num, done, h_taps); }
    return SDR_OK;
}

}  // namespace sdr
