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
    __shared__ uint32_t s_lut[41];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *stage = s_stage[warp];
    // plain variant: the pixel word is a function of the iteration count alone,
    // toGreen . truncate $ fromIntegral iCnt / 40 * 255  (Fractal2D.hs:93-97)
    if (threadIdx.x < 41) s_lut[threadIdx.x] = ((uint32_t)(int)truncf((float)threadIdx.x / 40.0f * 255.0f)) << 8;
    __syncthreads();
    const int mx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
    const int my = ((lane >> 1) & 1) | ((lane >> 2) & 2);
    const int k = P.resolve_log2;
    const float log_2 = rg_log(2.0f);
    unsigned iters = 0, samples = 0;

    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(P.queue, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= (unsigned)P.n_items) break;
        // item → (band of this rank, block row inside the band, block column)
        const int per_band = P.blocks_x * P.blocks_y;
        const int band_k = (int)item / per_band, in_band = (int)item - band_k * per_band;
        const int by = in_band / P.blocks_x, bx = in_band - by * P.blocks_x;
        const int band = band_k * P.shard_count + P.shard_rank;
        const int bw = min(kBlockDim, P.w - bx * kBlockDim), bh = min(P.item_h, P.band_h - by * P.item_h);
        const int X0 = bx * kBlockDim, Y0 = P.y0 + band * P.band_h + by * P.item_h;

        // x depends on the column only (Fractal2D.hs:82): four columns per lane and block
        float xs[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            xs[c] = ((float)(X0 + c * 8 + mx) / P.fw) * 2.9f * P.ratio - P.xshift;
            asm volatile("" : "+f"(xs[c]));          // keep the value: without this ptxas re-does the IEEE division per sample
        }

#pragma unroll 1
        for (int trow = 0; trow < (P.item_h >> 2); trow++) {
            const int ty = trow * 4;
            if (ty >= bh) break;
            const int ly = ty + my;
            const float y = ((float)(Y0 + ly) / P.fh) * 2.9f - 1.45f;               // :81
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int lx = c * 8 + mx;
                if (c * 8 >= bw) break;                                             // warp-uniform
                if (lx < bw && ly < bh) {
                    float zr = xs[c], zi = y;
                    float zr2 = zr * zr, zi2 = zi * zi;
                    int iter = 40;
                    // go, Fractal2D.hs:85-90, five blocks of eight iterations.  The reference's
                    // 1-cycle test (newZ == z ⇒ maxIter) is only a shortcut: a repeating orbit never
                    // escapes, so it ends at maxIter with the same z either way.  Testing once per block
                    // for z == z(block start) catches every cycle whose period divides 8 — same result
                    // (maxIter), one compare pair per eight iterations instead of per iteration.
#pragma unroll 1
                    for (int blk = 0; blk < 5; blk++) {
                        const float sr0 = zr, si0 = zi;
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            if ((zr2 + zi2) > 4.0f * 4.0f) { iter = blk * 8 + u; blk = 5; break; }
                            const float nr = (zr2 - zi2) + P.julia_r;
                            const float ni = (zr * zi + zi * zr) + P.julia_i;
                            zr = nr; zi = ni;
                            zr2 = zr * zr; zi2 = zi * zi;
                        }
                        if (blk < 5 && zr == sr0 && zi == si0) break;                // periodic: interior
                    }
                    iters += (unsigned)iter;
                    uint32_t word;
                    if (P.smooth && iter != 40) {
                        const float v = (float)iter - rg_log(rg_log(zr2 + zi2)) / log_2;
                        const float ic = v > 0.0f ? v : 0.0f;
                        word = ((uint32_t)(int)truncf(ic / 40.0f * 255.0f)) << 8;
                    } else {
                        word = s_lut[iter];
                    }
                    stage[ly * kBlockDim + lx] = word;                               // toGreen, :93-94
                    samples++;
                }
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
        // packed: this rank's bands back to back; frame: every band at its place in the whole frame
        const long long o0 = ((long long)(P.frame_layout ? band : band_k) * (P.band_h >> k) + ((by * P.item_h) >> k)) * P.out_stride + (X0 >> k);
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
