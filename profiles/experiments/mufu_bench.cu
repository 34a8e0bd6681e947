// mufu_bench.cu — what does a MUFU (XU pipe) instruction cost next to FP32 work on sm_100a?
// The march kernels sit at ~70 % FMA-pipe utilisation with math_pipe_throttle as the top stall although no
// single pipe is full by ncu's accounting.  This measures (a) pure MUFU throughput per function and
// (b) FFMA : MUFU mixes at the ratios the Mandelbulb iteration has (≈ 40 FP32 : 2 MUFU).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ float mufu_rsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_lg2(float x) { float r; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_sqrt(float x) { float r; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_sin(float x) { float r; asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int FN> __device__ __forceinline__ float mufu(float x)
{
    if (FN == 0) return mufu_rsq(x);
    if (FN == 1) return mufu_rcp(x);
    if (FN == 2) return mufu_ex2(x);
    if (FN == 3) return mufu_lg2(x);
    if (FN == 4) return mufu_sqrt(x);
    return mufu_sin(x);
}

// NF scalar FFMA (8 independent chains) and NM MUFU (4 independent chains) per inner step
template <int FN, int NF, int NM>
__global__ void __launch_bounds__(256) k(float *sink, int iters)
{
    float a[8], f[4];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = threadIdx.x * 1e-3f + j;
#pragma unroll
    for (int j = 0; j < 4; j++) f[j] = 1.0f + threadIdx.x * 1e-3f + j;
    const float m = 0.999f, c = 1e-3f;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int j = 0; j < NF; j++) a[j & 7] = fmaf(a[j & 7], m, c);
#pragma unroll
            for (int j = 0; j < NM; j++) f[j & 3] = mufu<FN>(f[j & 3]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
#pragma unroll
    for (int j = 0; j < 4; j++) s += f[j];
    if (s == 123.456f) sink[0] = s;
}

template <int FN, int NF, int NM>
static void run(const char *name, int sms, float *sink, double clk_hz)
{
    const int iters = 2048, threads = 256, warps_per_sm = 32, grid = sms * warps_per_sm * 32 / threads;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<FN, NF, NM><<<grid, threads>>>(sink, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 1 && ms < best) best = ms;
    }
    // cycles per inner step per SM sub-partition: 8 warps per sub-partition run `iters*8` steps each
    const double cyc = best * 1e-3 * clk_hz / (iters * 8.0) / (warps_per_sm / 4.0);
    printf("%-28s %2d FFMA + %d MUFU: %7.3f ms  %6.2f cycles/step/warp on its sub-partition  (FFMA alone = %d, MUFU at 8 cyc = %d)\n",
           name, NF, NM, best, cyc, NF, NM * 8);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double hz = clk * 1e3;
    float *sink; cudaMalloc(&sink, 4);
    const int n = p.multiProcessorCount;
    run<0, 0, 4>("rsqrt", n, sink, hz);  run<1, 0, 4>("rcp", n, sink, hz);  run<2, 0, 4>("ex2", n, sink, hz);
    run<3, 0, 4>("lg2", n, sink, hz);    run<4, 0, 4>("sqrt", n, sink, hz); run<5, 0, 4>("sin", n, sink, hz);
    run<0, 16, 0>("ffma only", n, sink, hz);
    run<0, 40, 2>("rsqrt", n, sink, hz); run<0, 40, 1>("rsqrt", n, sink, hz); run<0, 40, 4>("rsqrt", n, sink, hz);
    run<0, 20, 2>("rsqrt", n, sink, hz); run<0, 16, 2>("rsqrt", n, sink, hz); run<0, 8, 2>("rsqrt", n, sink, hz);
    run<2, 40, 2>("ex2", n, sink, hz);   run<2, 16, 2>("ex2", n, sink, hz);   run<1, 40, 2>("rcp", n, sink, hz);
    run<3, 40, 2>("lg2", n, sink, hz);   run<4, 40, 2>("sqrt", n, sink, hz);
    return 0;
}
