// fm_timeline.cu -- per-warp time line of one launch of the fused FM front end (measurement aid, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include tools/fm_timeline.cu -o build/fm_timeline
//   build/fm_timeline [log2 samples]
// stamps (globaltimer): 0 entry, 1 after barrier init + tap loads + griddepcontrol.wait, 2 after the first fills are issued,
// 3 first sub-tile's data has landed, 4 end of the main loop, 5 end of the boundary pass.
#define SDR_FM_TIMING 1
#include "../sdr_b200/csrc/kernels_fm.cu"

#include <algorithm>
#include <vector>

#include <cstdarg>

namespace sdr {   // the three helpers the launcher takes from ctx.cu, stand-alone
int set_error(int code, const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return code; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    fprintf(stderr, "%s:%d %s: %s\n", file, line, what, cudaGetErrorString(e)); return SDR_ECUDA;
}
int ring_attr(Ctx *, const void *kernel, int smem_bytes) {
    SDR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); return SDR_OK;
}
int Ctx::bind() const { return SDR_OK; }
}
using namespace sdr;

int main(int argc, char **argv) {
    const int log2 = argc > 1 ? atoi(argv[1]) : 24;
    const long long n = 1LL << log2;               // IQ pairs
    Ctx c;
    c.device = 0; cudaSetDevice(0);
    cudaStreamCreate(&c.stream);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0); c.sm_count = prop.multiProcessorCount;
    uint8_t *d_in; float *d_out, *d_taps; float2 *d_bnd, *d_carry; unsigned int *d_ticket;
    cudaMalloc(&d_in, 2 * n + 256); cudaMalloc(&d_out, 4 * (n / 8) + 256); cudaMalloc(&d_taps, 4 * 256);
    cudaMalloc(&d_bnd, 16 * (n / 8 / 256 + 2)); cudaMalloc(&d_carry, 32); cudaMalloc(&d_ticket, 4);
    cudaMemset(d_carry, 0, 32); cudaMemset(d_ticket, 0, 4);
    std::vector<uint8_t> h(2 * n); for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 13);
    cudaMemcpy(d_in, h.data(), h.size(), cudaMemcpyHostToDevice);
    std::vector<float> taps(256, 0.0f); for (int k = 0; k < 64; k++) taps[k] = taps[127 - k] = 0.01f * (k + 1);
    cudaMemcpy(d_taps, taps.data(), 4 * 256, cudaMemcpyHostToDevice);
    const long long num = (n - 128) / 8 + 1;
    long long done = 0; const char *name = nullptr;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, c.stream);
        int st = launch_fm_front(&c, 128, 8, d_taps, true, d_in, n, nullptr, n, d_out, num, d_bnd, n / 8 / 256 + 2, d_carry, d_carry + 1, d_ticket, &done, &name);
        cudaEventRecord(e1, c.stream);
        cudaStreamSynchronize(c.stream);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d status %d %s done %lld: %.1f us (events)\n", rep, st, name, done, ms * 1e3);
    }
    for (int rep = 0; rep < 3; rep++) {   // cadence of back-to-back launches (programmatic dependent launch unless SDR_B200_NO_PDL)
        cudaEventRecord(e0, c.stream);
        for (int k = 0; k < 16; k++)
            launch_fm_front(&c, 128, 8, d_taps, true, d_in, n, nullptr, n, d_out, num, d_bnd, n / 8 / 256 + 2, d_carry, d_carry + 1, d_ticket, &done, &name);
        cudaEventRecord(e1, c.stream);
        cudaStreamSynchronize(c.stream);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("16 launches back to back: %.1f us each\n", ms * 1e3 / 16);
    }
    std::vector<unsigned long long> t(160 * 16 * 8);
    cudaMemcpyFromSymbol(t.data(), g_fm_timing, t.size() * 8);
    unsigned long long t0 = ~0ULL;
    for (int b = 0; b < c.sm_count; b++) for (int w = 0; w < 16; w++) t0 = std::min(t0, t[(b * 16 + w) * 8]);
    const char *lab[6] = {"entry", "after init+taps+wait", "first fills issued", "first data landed", "main loop done", "boundary pass done"};
    for (int i = 0; i < 6; i++) {
        std::vector<double> v;
        for (int b = 0; b < c.sm_count; b++) for (int w = 0; w < 16; w++) v.push_back((double)(t[(b * 16 + w) * 8 + i] - t0) * 1e-3);
        std::sort(v.begin(), v.end());
        printf("%-22s min %7.2f  median %7.2f  p90 %7.2f  max %7.2f us\n", lab[i], v.front(), v[v.size() / 2], v[v.size() * 9 / 10], v.back());
    }
    // per-warp main-loop duration by number of sub-tiles
    const long long n_sub = (num + 255) / 256;
    printf("sub-tiles %lld = %.2f per CTA, %.2f per warp\n", n_sub, (double)n_sub / c.sm_count, (double)n_sub / c.sm_count / 16);
    for (int w = 0; w < 16; w++) {
        double s = 0; for (int b = 0; b < c.sm_count; b++) s += (double)(t[(b * 16 + w) * 8 + 4] - t[(b * 16 + w) * 8 + 3]) * 1e-3;
        printf("warp %2d: main loop %.2f us (mean over CTAs)\n", w, s / c.sm_count);
    }
    return 0;
}
