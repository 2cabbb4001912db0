// mb_fma.cu -- FP32 pipe micro-benchmark for sm_100a (measurement tool, not product code).
// Answers: how many lane-FMAs per clock per SM do (a) 3-register FFMA, (b) FFMA with a constant-bank operand,
// (c) packed FFMA2 (fma.rn.f32x2) with duplicated-coefficient register pairs, (d) FFMA2 interleaved with LDS.128
// sustain?  Drives the choice of inner loop for the decimator (DESIGN.md "FP32 issue budget").
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_fma mb_fma.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d;
}
__device__ __forceinline__ float2 unpack2(u64 v) {
    float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r;
}
__constant__ float ctaps[128];

#define NACC 32
// (a) 3-register FFMA: 32 accumulators, 8 x, 8 c
__global__ void __launch_bounds__(256) k_rrr(const float *in, float *out, int iters, long long *clk) {
    float acc[NACC], x[8], c[8];
    for (int i = 0; i < NACC; i++) acc[i] = in[threadIdx.x + i];
    for (int i = 0; i < 8; i++) { x[i] = in[64 + threadIdx.x + i]; c[i] = in[128 + threadIdx.x + i]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = fmaf(x[(i + j) & 7], c[j], acc[i]);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
// (b) constant-bank operand
__global__ void __launch_bounds__(256) k_const(const float *in, float *out, int iters, long long *clk) {
    float acc[NACC], x[8];
    for (int i = 0; i < NACC; i++) acc[i] = in[threadIdx.x + i];
    for (int i = 0; i < 8; i++) x[i] = in[64 + threadIdx.x + i];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = fmaf(x[(i + j) & 7], ctaps[(j * 16 + i) & 127], acc[i]);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
// (c) FFMA2, coefficient pairs pre-duplicated in registers: 16 acc pairs, 8 x pairs, 8 c pairs
__global__ void __launch_bounds__(256) k_ffma2(const float *in, float *out, int iters, long long *clk) {
    u64 acc[16], x[8], c[8];
    for (int i = 0; i < 16; i++) acc[i] = pack2(in[threadIdx.x + i], in[threadIdx.x + i + 16]);
    for (int i = 0; i < 8; i++) { x[i] = pack2(in[64 + threadIdx.x + i], in[80 + threadIdx.x + i]);
                                  float cc = in[128 + threadIdx.x + i]; c[i] = pack2(cc, cc); }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++)
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = ffma2(x[(i + j) & 7], c[j & 7], acc[i]);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; i++) { float2 v = unpack2(acc[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
// (d) FFMA2 with one LDS.128 per 16 FFMA2 (x pairs refreshed from shared memory) + one broadcast LDS.64 per 16
__global__ void __launch_bounds__(256) k_ffma2_lds(const float *in, float *out, int iters, long long *clk) {
    __shared__ float4 sm[256 * 2 + 64];
    for (int i = threadIdx.x; i < 256 * 2 + 64; i += 256) sm[i] = make_float4(in[i & 255], in[(i + 1) & 255], 1.f, 2.f);
    __syncthreads();
    u64 acc[16], x[8], c[8];
    for (int i = 0; i < 16; i++) acc[i] = pack2(in[threadIdx.x + i], in[threadIdx.x + i + 16]);
    for (int i = 0; i < 8; i++) { x[i] = pack2(in[64 + threadIdx.x + i], in[80 + threadIdx.x + i]);
                                  float cc = in[128 + threadIdx.x + i]; c[i] = pack2(cc, cc); }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            float4 v = sm[threadIdx.x + ((it + j) & 63) * 4];              // conflict-free LDS.128
            x[(2 * j) & 7] = pack2(v.x, v.y); x[(2 * j + 1) & 7] = pack2(v.z, v.w);
            u64 cv = *reinterpret_cast<const u64 *>(&sm[512 + ((it + j) & 63)]);   // broadcast LDS.64
            c[j & 7] = cv;
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = ffma2(x[(i + j) & 7], c[(i + j) & 7], acc[i]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; i++) { float2 v = unpack2(acc[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
// (e) scalar FFMA const-bank with one LDS.128 per 64 FFMA
__global__ void __launch_bounds__(256) k_const_lds(const float *in, float *out, int iters, long long *clk) {
    __shared__ float4 sm[256 * 2 + 64];
    for (int i = threadIdx.x; i < 256 * 2 + 64; i += 256) sm[i] = make_float4(in[i & 255], in[(i + 1) & 255], 1.f, 2.f);
    __syncthreads();
    float acc[NACC], x[8];
    for (int i = 0; i < NACC; i++) acc[i] = in[threadIdx.x + i];
    for (int i = 0; i < 8; i++) x[i] = in[64 + threadIdx.x + i];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if ((j & 1) == 0) { float4 v = sm[threadIdx.x + ((it + j) & 63) * 4];
                                x[(2 * j) & 7] = v.x; x[(2 * j + 1) & 7] = v.y; x[(2 * j + 2) & 7] = v.z; x[(2 * j + 3) & 7] = v.w; }
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = fmaf(x[(i + j) & 7], ctaps[(j * 16 + i) & 127], acc[i]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

typedef void (*kern_t)(const float *, float *, int, long long *);
struct Case { const char *name; kern_t k; double fma_per_thread_iter; };

int main() {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int sms = p.multiProcessorCount;
    float h[512]; for (int i = 0; i < 512; i++) h[i] = 1.0f / (1 + i);
    float *din, *dout; long long *dclk;
    cudaMalloc(&din, sizeof(h)); cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(ctaps, h, 128 * 4);
    cudaMalloc(&dout, 148 * 8 * 256 * 4 * 4); cudaMalloc(&dclk, 148 * 8 * 8 * 8);
    Case cases[] = { {"ffma_rrr", k_rrr, 8.0 * NACC}, {"ffma_const", k_const, 8.0 * NACC},
                     {"ffma2_dup", k_ffma2, 2.0 * 256}, {"ffma2_dup+lds", k_ffma2_lds, 2.0 * 256},
                     {"ffma_const+lds", k_const_lds, 8.0 * NACC} };
    printf("device %s, %d SMs\n", p.name, sms);
    for (auto &cs : cases) {
        for (int ctas_per_sm = 1; ctas_per_sm <= 2; ctas_per_sm++) {
            int iters = 20000, grid = sms * ctas_per_sm;
            cs.k<<<grid, 256>>>(din, dout, 100, dclk);   // warm
            cudaDeviceSynchronize();
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            cs.k<<<grid, 256>>>(din, dout, iters, dclk);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long hc[148 * 2]; cudaMemcpy(hc, dclk, grid * 8, cudaMemcpyDeviceToHost);
            double avgclk = 0; for (int i = 0; i < grid; i++) avgclk += hc[i]; avgclk /= grid;
            double fma_sm = cs.fma_per_thread_iter * iters * 256.0 * ctas_per_sm;   // lane-FMAs per SM
            double mhz = avgclk / (ms * 1e3);
            printf("%-16s warps/SM %2d : %.3f ms, %.0f clk, ~%.0f MHz, %.1f lane-FMA/clk/SM, %.2f TFLOP/s (err=%s)\n",
                   cs.name, 8 * ctas_per_sm, ms, avgclk, mhz, fma_sm / avgclk,
                   2.0 * fma_sm * sms / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
