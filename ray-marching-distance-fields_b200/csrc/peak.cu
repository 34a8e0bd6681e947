// peak.cu — FP32 FMA roofline denominator, measured in the same process as the benchmark
// (MEASURED_PEAKS.json carries HBM and bf16-tensor figures only).  Every thread runs 8
// independent FFMA chains; 2 flop per FFMA per lane.  The SM clock actually sustained is
// taken from clock64() so the result can be compared with 148 SM × 128 lanes × 2 × f.
#include <cuda_runtime.h>
#include <stdint.h>
#include "rmdf_internal.h"

namespace rmdf {

__global__ void __launch_bounds__(256)
ffma_peak_kernel(float *sink, int iters, long long *cycles)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456f) sink[0] = s;
}

cudaError_t measure_fp32_peak(int sm_count, cudaStream_t s, float *tflops, float *sm_mhz)
{
    float *sink = nullptr; long long *cyc = nullptr;
    cudaError_t e = cudaMalloc(&sink, 4); if (e != cudaSuccess) return e;
    e = cudaMalloc(&cyc, 8); if (e != cudaSuccess) { cudaFree(sink); return e; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sm_count * 8, iters = 4096;
    float best_ms = 1e30f; long long cycles = 0;
    for (int rep = 0; rep < 6; rep++) {               // first reps warm the clocks up
        cudaEventRecord(e0, s);
        ffma_peak_kernel<<<grid, 256, 0, s>>>(sink, iters, cyc);
        cudaEventRecord(e1, s);
        e = cudaStreamSynchronize(s); if (e != cudaSuccess) break;
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best_ms) { best_ms = ms; cudaMemcpy(&cycles, cyc, 8, cudaMemcpyDeviceToHost); }
    }
    if (e == cudaSuccess) {
        const double flop = 2.0 * 8.0 * 16.0 * (double)iters * 256.0 * (double)grid;
        *tflops = (float)(flop / (best_ms * 1e-3) / 1e12);
        // one CTA's loop spans ~the whole kernel when 8 CTAs/SM are co-resident
        *sm_mhz = (float)((double)cycles / (best_ms * 1e-3) / 1e6);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink); cudaFree(cyc);
    return e;
}

// MUFU (XU pipe) roofline denominator for the transcendental-bound scenes (general-power Mandelbulb,
// CSG test scene): four independent ex2.approx chains per thread.  sm_100a issues one MUFU warp
// instruction per 8 cycles per SM sub-partition = 16 lanes/clk/SM (profiles/experiments/mufu_bench.cu:
// rsq, ex2, lg2, sqrt and sin all measure the same 8.04 cycles).
__global__ void __launch_bounds__(256)
mufu_peak_kernel(float *sink, int iters)
{
    float f0 = 1.0f + threadIdx.x * 1e-3f, f1 = f0 + 1.f, f2 = f0 + 2.f, f3 = f0 + 3.f;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f3));
        }
    }
    const float s = (f0 + f1) + (f2 + f3);
    if (s == 123.456f) sink[0] = s;
}

cudaError_t measure_mufu_peak(int sm_count, cudaStream_t s, float *gops)
{
    float *sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 4); if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sm_count * 8, iters = 1024;
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, s);
        mufu_peak_kernel<<<grid, 256, 0, s>>>(sink, iters);
        cudaEventRecord(e1, s);
        e = cudaStreamSynchronize(s); if (e != cudaSuccess) break;
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best_ms) best_ms = ms;
    }
    if (e == cudaSuccess) *gops = (float)(4.0 * 16.0 * (double)iters * 256.0 * (double)grid / (best_ms * 1e-3) / 1e9);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return e;
}

}  // namespace rmdf
