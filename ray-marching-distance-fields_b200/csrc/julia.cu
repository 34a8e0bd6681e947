// julia.cu — the animated 2D Julia set of /root/reference/Fractal2D.hs:63-98
// (juliaAnimated; magnitudeSq :19-20, fractionalIterCnt :24-25), which the reference
// computes on the CPU with forSegmentsConcurrently and uploads through fillFrameBuffer.
//
// Compiled -fmad=false: the orbit z ← z² + c uses individually rounded IEEE mul/add in
// the order of Data.Complex's (*) and (+), so iteration counts — and therefore the
// integer green-channel output — are bit-identical to the CPU oracle.  The smooth
// variant's log(log |z|²)/log 2 uses the deterministic rg_log (≤ 1 ulp from libm).
//
// Same work decomposition as the march kernel: persistent CTAs, one atomicAdd per
// 32×32 block of samples, 8×4 Morton tiles per warp, RGBA8 staging in shared memory,
// in-place mip resolve (supersampling), 128-bit stores.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rmdf_glsl.h"
#include "rmdf_internal.h"

namespace rmdf {

__device__ __forceinline__ uint32_t j_box4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    uint32_t lo = (a & 0x00ff00ffu) + (b & 0x00ff00ffu) + (c & 0x00ff00ffu) + (d & 0x00ff00ffu) + 0x00020002u;
    uint32_t hi = ((a >> 8) & 0x00ff00ffu) + ((b >> 8) & 0x00ff00ffu) + ((c >> 8) & 0x00ff00ffu) + ((d >> 8) & 0x00ff00ffu) + 0x00020002u;
    return ((lo >> 2) & 0x00ff00ffu) | (((hi >> 2) & 0x00ff00ffu) << 8);
}

__global__ void __launch_bounds__(kCtaThreads, 8)
julia_kernel(const JuliaParams P)
{
    __shared__ __align__(16) uint32_t s_stage[kWarpsPerCta][kBlockDim * kBlockDim];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *stage = s_stage[warp];
    const int mx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
    const int my = ((lane >> 1) & 1) | ((lane >> 2) & 2);
    const int k = P.resolve_log2;
    const float log2_inv_den = rg_log(2.0f);
    unsigned iters = 0, samples = 0;

    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(P.queue, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= (unsigned)P.n_items) break;
        const int by = (int)item / P.blocks_x, bx = (int)item - by * P.blocks_x;
        const int rh = P.y1 - P.y0;
        const int bw = min(kBlockDim, P.w - bx * kBlockDim), bh = min(P.item_h, rh - by * P.item_h);
        const int X0 = bx * kBlockDim, Y0 = P.y0 + by * P.item_h;

#pragma unroll 1
        for (int tile = 0; tile < P.item_h; tile++) {
            const int tx = (tile & 3) * 8, ty = (tile >> 2) * 4;
            if (tx >= bw || ty >= bh) continue;
            const int lx = tx + mx, ly = ty + my;
            if (lx < bw && ly < bh) {
                const float fpx = (float)(X0 + lx), fpy = (float)(Y0 + ly);
                const float y = (fpy / P.fh) * 2.9f - 1.45f;                 // :81
                const float x = (fpx / P.fw) * 2.9f * P.ratio - P.xshift;    // :82
                float zr = x, zi = y;
                int iter = 0;
                for (;;) {                                                    // go, :85-90
                    if (iter == 40 || (zr * zr + zi * zi) > 4.0f * 4.0f) break;
                    const float nr = (zr * zr - zi * zi) + P.julia_r;
                    const float ni = (zr * zi + zi * zr) + P.julia_i;
                    if (nr == zr && ni == zi) { iter = 40; break; }           // 1-cycle detection
                    zr = nr; zi = ni;
                    iter++;
                    iters++;
                }
                float val;
                if (P.smooth) {
                    float ic = 40.0f;
                    if (iter != 40) {
                        const float v = (float)iter - rg_log(rg_log(zr * zr + zi * zi)) / log2_inv_den;
                        ic = v > 0.0f ? v : 0.0f;
                    }
                    val = ic / 40.0f * 255.0f;
                } else {
                    val = (float)iter / 40.0f * 255.0f;
                }
                stage[ly * kBlockDim + lx] = ((uint32_t)(int)truncf(val)) << 8;   // toGreen, :93-94
                samples++;
            }
        }
        __syncwarp();
        for (int l = 1; l <= k; l++) {
            const int lw = bw >> l, lh = bh >> l, h = 1 << (l - 1);
            for (int g = lane; g < lw * lh; g += 32) {
                const int y = g / lw, x = g - y * lw;
                uint32_t *p = stage + ((y << l) * kBlockDim + (x << l));
                p[0] = j_box4(p[0], p[h], p[h * kBlockDim], p[h * kBlockDim + h]);
            }
            __syncwarp();
        }
        const int ow = bw >> k, oh = bh >> k;
        const long long o0 = (long long)((Y0 - P.y0) >> k) * P.out_stride + (X0 >> k);
        const bool vec_ok = ((ow & 3) == 0) && ((P.out_stride & 3) == 0) && ((o0 & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0);
        if (vec_ok) {
            const int groups = (ow >> 2) * oh;
            for (int g = lane; g < groups; g += 32) {
                const int oy = g / (ow >> 2), ox = (g - oy * (ow >> 2)) << 2;
                const uint32_t *src = stage + ((oy << k) * kBlockDim + (ox << k));
                uint4 v;
                if (k == 0) v = *reinterpret_cast<const uint4 *>(src);
                else { v.x = src[0]; v.y = src[1 << k]; v.z = src[2 << k]; v.w = src[3 << k]; }
                *reinterpret_cast<uint4 *>(P.out + o0 + (long long)oy * P.out_stride + ox) = v;
            }
        } else {
            for (int g = lane; g < ow * oh; g += 32) {
                const int oy = g / ow, ox = g - oy * ow;
                P.out[o0 + (long long)oy * P.out_stride + ox] = stage[(oy << k) * kBlockDim + (ox << k)];
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        iters += __shfl_xor_sync(0xffffffffu, iters, o);
        samples += __shfl_xor_sync(0xffffffffu, samples, o);
    }
    if (lane == 0) {
        atomicAdd(P.counters + C_SAMPLES, (unsigned long long)samples);
        atomicAdd(P.counters + C_ITERS, (unsigned long long)iters);
    }
}

cudaError_t launch_julia(const JuliaParams &p, int grid, cudaStream_t s)
{
    julia_kernel<<<grid, kCtaThreads, 0, s>>>(p);
    return cudaGetLastError();
}
int julia_ctas_per_sm()
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, julia_kernel, kCtaThreads, 0) != cudaSuccess || n < 1) n = 1;
    return n;
}

}  // namespace rmdf
