// ffma2_bench.cu — does sm_100a's packed FP32 (FFMA2 / FMUL2 / FADD2, PTX fma.rn.f32x2) free
// issue slots?  The march kernels are instruction-issue bound (profiles/r1_march_fast_final.md:
// issue 94.7 %, FMA pipe 68 %), so what matters is (a) FP32 lanes/clk/SM of the packed forms and
// (b) whether other pipes (ALU, XU) issue in the slots the packed forms leave free.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cuda_runtime.h>
#include <stdio.h>

#define UNROLL 16

template <int MODE>
__global__ void __launch_bounds__(256) k(float *sink, int iters)
{
    float2 a[8];
    unsigned u0 = threadIdx.x, u1 = threadIdx.x * 3u, u2 = 7u, u3 = 11u;
    float f0 = 1.0f + threadIdx.x, f1 = 2.0f + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = make_float2(threadIdx.x * 1e-3f + j, threadIdx.x * 2e-3f + j);
    const float2 m = make_float2(0.999f, 0.998f), c = make_float2(1e-3f, 2e-3f);
    float2 b[8], c2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { b[j] = make_float2(0.999f - 1e-4f * j - 1e-6f * threadIdx.x, 0.998f - 1e-4f * j); c2[j] = make_float2(1e-3f * j + 1e-6f * threadIdx.x, 2e-3f * j); }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (MODE == 0) {            // 16 scalar FFMA
#pragma unroll
                for (int j = 0; j < 8; j++) { a[j].x = fmaf(a[j].x, m.x, c.x); a[j].y = fmaf(a[j].y, m.y, c.y); }
            } else if (MODE == 1) {     // 8 FFMA2
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = __ffma2_rn(a[j], m, c);
            } else if (MODE == 2) {     // 8 FMUL2
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = __fmul2_rn(a[j], m);
            } else if (MODE == 3) {     // 8 FADD2
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = __fadd2_rn(a[j], c);
            } else if (MODE == 4) {     // 8 FFMA2 + 8 ALU (LOP3/IADD3)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    a[j] = __ffma2_rn(a[j], m, c);
                    if (j & 1) u0 = (u0 ^ u1) + u2; else u1 = (u1 & u3) ^ u0;
                }
            } else if (MODE == 5) {     // 16 scalar FFMA + 8 ALU
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    a[j].x = fmaf(a[j].x, m.x, c.x); a[j].y = fmaf(a[j].y, m.y, c.y);
                    if (j & 1) u0 = (u0 ^ u1) + u2; else u1 = (u1 & u3) ^ u0;
                }
            } else if (MODE == 6) {     // 8 FFMA2 + 2 MUFU.RSQ
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = __ffma2_rn(a[j], m, c);
                f0 = rsqrtf(f0); f1 = rsqrtf(f1);
            } else if (MODE == 7) {     // 16 scalar FFMA + 2 MUFU.RSQ
#pragma unroll
                for (int j = 0; j < 8; j++) { a[j].x = fmaf(a[j].x, m.x, c.x); a[j].y = fmaf(a[j].y, m.y, c.y); }
                f0 = rsqrtf(f0); f1 = rsqrtf(f1);
            } else if (MODE == 8) {     // 8 FFMA2 + 4 MUFU + 8 ALU   (≈ the packed Mandelbulb mix)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    a[j] = __ffma2_rn(a[j], m, c);
                    if (j & 1) u0 = (u0 ^ u1) + u2; else u1 = (u1 & u3) ^ u0;
                }
                f0 = rsqrtf(f0); f1 = rsqrtf(f1); f0 = rsqrtf(f0); f1 = rsqrtf(f1);
            } else if (MODE == 9) {     // 4 FFMA2 + 8 scalar FFMA (mixed packing)
#pragma unroll
                for (int j = 0; j < 4; j++) a[j] = __ffma2_rn(a[j], m, c);
#pragma unroll
                for (int j = 4; j < 8; j++) { a[j].x = fmaf(a[j].x, m.x, c.x); a[j].y = fmaf(a[j].y, m.y, c.y); }
            } else if (MODE == 12) {    // 8 FFMA2, three distinct register-pair operands each
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = __ffma2_rn(a[j], b[j], c2[j]);
            } else if (MODE == 13) {    // 16 FFMA, three distinct register operands each
#pragma unroll
                for (int j = 0; j < 8; j++) { a[j].x = fmaf(a[j].x, b[j].x, c2[j].x); a[j].y = fmaf(a[j].y, b[j].y, c2[j].y); }
            } else if (MODE == 14) {    // 8 FMUL2, two distinct register-pair operands
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = __fmul2_rn(a[j], b[j]);
            } else if (MODE == 15) {    // 8 FFMA2 a = a*b + a (two distinct pairs)
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = __ffma2_rn(a[j], b[j], a[j]);
            } else if (MODE == 10) {    // dependent chain latency: 1 FFMA2 chain
                a[0] = __ffma2_rn(a[0], m, c);
            } else if (MODE == 11) {    // dependent chain latency: 1 FFMA chain
                a[0].x = fmaf(a[0].x, m.x, c.x);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j].x + a[j].y;
    s += (float)(u0 + u1) + f0 + f1;
#pragma unroll
    for (int j = 0; j < 8; j++) s += b[j].x + b[j].y + c2[j].x + c2[j].y;
    if (s == 123.456f) sink[0] = s;
}

template <int MODE>
static void run(const char *name, double fp32_per_thread_iter, int sms, float *sink, int warps_per_sm)
{
    const int iters = 2048;
    const int threads = 256, grid = sms * warps_per_sm * 32 / threads;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<grid, threads>>>(sink, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 1 && ms < best) best = ms;
    }
    const double lanes_ops = fp32_per_thread_iter * UNROLL * (double)iters * threads * grid;   // FP32 lane-operations
    printf("%-44s warps/SM %2d  %8.3f ms  %7.2f T lane-op/s  (x2 flop for FMA)\n", name, warps_per_sm, best, lanes_ops / (best * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, %d kHz: 128 lanes x SMs x clk = %.2f T lane-op/s\n", p.name, p.multiProcessorCount, clk,
           128.0 * p.multiProcessorCount * clk * 1e3 / 1e12);
    float *sink; cudaMalloc(&sink, 4);
    for (int w : {16, 32}) {
        run<0>("16 FFMA", 16, p.multiProcessorCount, sink, w);
        run<1>("8 FFMA2", 16, p.multiProcessorCount, sink, w);
        run<2>("8 FMUL2", 16, p.multiProcessorCount, sink, w);
        run<3>("8 FADD2", 16, p.multiProcessorCount, sink, w);
        run<4>("8 FFMA2 + 8 ALU", 16, p.multiProcessorCount, sink, w);
        run<5>("16 FFMA + 8 ALU", 16, p.multiProcessorCount, sink, w);
        run<6>("8 FFMA2 + 2 MUFU", 16, p.multiProcessorCount, sink, w);
        run<7>("16 FFMA + 2 MUFU", 16, p.multiProcessorCount, sink, w);
        run<8>("8 FFMA2 + 4 MUFU + 8 ALU", 16, p.multiProcessorCount, sink, w);
        run<9>("4 FFMA2 + 8 FFMA", 16, p.multiProcessorCount, sink, w);
        run<12>("8 FFMA2, 3 distinct reg pairs", 16, p.multiProcessorCount, sink, w);
        run<13>("16 FFMA, 3 distinct regs", 16, p.multiProcessorCount, sink, w);
        run<14>("8 FMUL2, 2 distinct reg pairs", 16, p.multiProcessorCount, sink, w);
        run<15>("8 FFMA2 a*b+a", 16, p.multiProcessorCount, sink, w);
    }
    run<10>("1 FFMA2 dependent chain (latency)", 2, p.multiProcessorCount, sink, 4);
    run<11>("1 FFMA dependent chain (latency)", 1, p.multiProcessorCount, sink, 4);
    return 0;
}
