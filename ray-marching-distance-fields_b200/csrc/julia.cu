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
// 32×item_h block of samples, RGBA8 staging in shared memory, in-place mip resolve
// (supersampling), 128-bit stores.  A warp covers 16×4 samples, two adjacent columns per
// lane, and iterates both in packed FP32 (csrc/packed_f32.cuh).
//
// The frame of BASELINE config 5 averages 3.8 iterations per sample, so the per-sample
// bookkeeping weighs as much as the orbit: the smooth switch is a template parameter, the
// first block of eight iterations is peeled (its periodicity reference is the start value,
// no copies), and the resolve works on the green byte alone (the other three are zero).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rmdf_glsl.h"
#include "rmdf_internal.h"
#include "packed_f32.cuh"

namespace rmdf {

// (a+b+c+d+2)>>2 per byte (the glGenerateMipmap box of oracle convention C6) for words that hold
// only a green byte, 0x0000GG00: the sum of four cannot carry out of bits 8..17.
__device__ __forceinline__ uint32_t j_box4_green(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return (((a + b) + (c + d) + 0x200u) >> 2) & 0xff00u;
}

// pixel word of one sample: toGreen . truncate $ fromIntegral iCnt / 40 * 255 (Fractal2D.hs:93-97); the smooth
// variant subtracts log(log |z|²)/log 2 at the escape (fractionalIterCnt, :24-25)
// a / b for a divisor known in advance, correctly rounded (Markstein): with y = RN(1/b), q = RN(a·y) is within one ulp
// of the quotient, r = a − b·q is exact in an FMA, and RN(q + r·y) is RN(a/b) for every b whose significand is not
// all ones — three instructions instead of the ~12 of the general IEEE division sequence.
__device__ __forceinline__ float div_known(float a, float b, float y)
{
    const float q = __fmul_rn(a, y);
    return __fmaf_rn(__fmaf_rn(-b, q, a), y, q);
}
template <bool SMOOTH>
__device__ __forceinline__ uint32_t julia_word(int iter, float mag2, float log_2, float rcp_log_2, uint32_t lut)
{
    if (SMOOTH && iter != 40) {
        // The byte is trunc(e) with e = max(iter − log2(ln |z|²), 0) · 255/40.  A MUFU evaluation of e (two lg2.approx) is
        // within ~3e-5 of the deterministic one below (|z|² in (16, 304], e <= 255: the lg2 errors and four roundings), so
        // wherever it lies at least 1e-3 away from an integer both truncate to the same byte: 99.8 % of the samples stop
        // here, the rest (and anything non-finite) take the exact sequence.
        {
            const float l1 = __log2f(mag2) * 0.693147181f;
            const float e = fmaxf((float)iter - __log2f(l1), 0.0f) * 6.375f;
            if (fabsf(e - rintf(e)) >= 1e-3f || e < 0.5f) return ((uint32_t)(int)e) << 8;
        }
        const float v = (float)iter - div_known(rg_log(rg_log(mag2)), log_2, rcp_log_2);
        const float ic = v > 0.0f ? v : 0.0f;
        return ((uint32_t)(int)truncf(div_known(ic, 40.0f, 0.025f) * 255.0f)) << 8;
    }
    // lut = shared-space address of the table: one LEA + LDS (indexing the __shared__ array re-derived its window base per use)
    uint32_t w;
    asm("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(lut + 4u * (uint32_t)iter));
    return w;
}

// Two orbit steps for both slots (go, Fractal2D.hs:85-90), with ONE escape test.  The reference tests |z|² > 16 before
// every update; here the test before the first update of a pair is skipped and made up for at the second: an orbit
// that has left the disc of radius 4 never comes back (|z² + c| >= |z|² − |c| > 4 for |c| < 2, and |c| <= 1.4 here),
// so it is still outside one step later — then the skipped test's |z|² (its operands are still in registers) says in
// which of the two iterations it left.  Nothing overflows before the test fires: |z| <= 4 at the last test that passed
// (or z is the start value, a coordinate of at most a few 10^5), so two updates stay finite.  A slot that has
// escaped is frozen by poisoning its additive constant with NaN: its orbit is NaN from the next update on and
// `|z|² > 16` stays false for it, so the common path costs one max + one compare for both slots per pair.
//
// Roundings are the reference's, operation by operation (Data.Complex: (zr·zr − zi·zi) + cr, (zr·zi + zi·zr) + ci):
//   2·p + ji as ONE packed FMA: p + p = 2p exactly, so RN(2p + ji) = RN((p + p) + ji);
//   zr² − zi² of the untested step as fma(zi², −1, zr²): exact product, so it is RN(zr² − zi²) — written this way because
//   ptxas contracts a packed product whose only use is a sum INTO that sum (packed_f32.cuh), which would round zr·zr − zi²
//   once instead of twice; as the addend of an explicit FMA there is nothing to contract.  Where both squares have a second
//   use (the tested step's |z|²) the plain difference is safe.  tests/test_host_cpu.py checks every FFMA2 of this unit
//   for its literal −1 / 2 multiplier.
#define RMDF_JULIA_UPDATE(diff)                                                                   \
    {                                                                                             \
        const f2 nr = (diff) + jr;                                                                \
        const f2 p = zr * zi;                                                                     \
        const f2 ni = fma2(2.0f, p, ji);                                                          \
        zr = nr; zi = ni;                                                                         \
        zr2 = zr * zr; zi2 = zi * zi;                                                             \
    }
#define RMDF_JULIA_PAIR(it)                                                                       \
    {                                                                                             \
        const f2 pr2 = zr2, pi2 = zi2;                               /* squares of z(it) */        \
        RMDF_JULIA_UPDATE(fma2(-1.0f, pi2, pr2))                                                  \
        const f2 m = zr2 + zi2;                                      /* |z(it + 1)|² */            \
        if (fmaxf(m.x, m.y) > 4.0f * 4.0f) {                                                      \
            const f2 m_prev = pr2 + pi2;                             /* the skipped test */        \
            if (m.x > 16.0f) {                                                                    \
                if (m_prev.x > 16.0f) { it0 = (it); m0 = m_prev.x; } else { it0 = (it) + 1; m0 = m.x; } \
                done |= 1u; jr.x = kNaN;                                                          \
            }                                                                                     \
            if (m.y > 16.0f) {                                                                    \
                if (m_prev.y > 16.0f) { it1 = (it); m1 = m_prev.y; } else { it1 = (it) + 1; m1 = m.y; } \
                done |= 2u; jr.y = kNaN;                                                          \
            }                                                                                     \
            if (done == 3u) goto finished;                                                        \
        }                                                                                         \
        RMDF_JULIA_UPDATE(zr2 - zi2)                                                              \
    }
// The reference's 1-cycle test (newZ == z ⇒ maxIter) is only a shortcut: a repeating orbit never escapes, so it
// ends at maxIter with the same z either way.  Testing once per block of eight iterations for
// z == z(block start) catches every cycle whose period divides the block length — same result (maxIter), one
// compare pair per block instead of per iteration.
#define RMDF_JULIA_PERIODIC(sr, si)                                                               \
    {                                                                                             \
        if (zr.x == (sr).x && zi.x == (si).x) { done |= 1u; jr.x = kNaN; }                        \
        if (zr.y == (sr).y && zi.y == (si).y) { done |= 2u; jr.y = kNaN; }                        \
        if (done == 3u) goto finished;                                                            \
    }

#ifndef RMDF_JULIA_BLOCK_LEN
#define RMDF_JULIA_BLOCK_LEN 8     // iterations between periodicity tests (a divisor of 40): 4 or 8
#endif
#if RMDF_JULIA_BLOCK_LEN == 8
#define RMDF_JULIA_BLOCK(b) RMDF_JULIA_PAIR(b) RMDF_JULIA_PAIR((b) + 2) RMDF_JULIA_PAIR((b) + 4) RMDF_JULIA_PAIR((b) + 6)
#else
#define RMDF_JULIA_BLOCK(b) RMDF_JULIA_PAIR(b) RMDF_JULIA_PAIR((b) + 2)
#endif

// KD: the coordinate divisions col / w and row / h (Fractal2D.hs:81-82) as correctly rounded known-divisor
// divisions (three instructions instead of the general IEEE sequence); the host sets it when neither divisor's
// significand is all ones (div_known's condition) — every frame size in practice.
template <bool SMOOTH, bool KD>
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
    const int mx = lane & 7, my = lane >> 3;        // a warp covers 16×4 samples: lane = 2 adjacent columns of one row
    const float kNaN = __int_as_float(0x7fffffff);
    uint32_t lut_addr = (uint32_t)__cvta_generic_to_shared(s_lut);
    asm volatile("" : "+r"(lut_addr));              // opaque: stays in a register (ptxas re-derived the shared window base, 4 instructions, at every use)
    const int k = P.resolve_log2;
    const float log_2 = rg_log(2.0f), rcp_log_2 = 1.0f / log_2;
    const f2 ji = F2(P.julia_i);
    unsigned iters = 0, samples = 0;

    // ---- one work item per trip (frame_sync.cuh: the first one by the warp's place in the grid, the others claimed singly) ----
    for (unsigned item = first_item<false>(P.sync, lane); item < (unsigned)P.n_items; item = claim_item(P.sync, lane)) {
        // item → (band of this rank, block row inside the band, block column)
        const int per_band = P.blocks_x * P.blocks_y;
        const unsigned fi = frame_item(P.sync, item);                                  // interleaved multi-GPU deal (frame_sync.cuh)
        if (fi >= P.sync.items_total) continue;
        const int fitem = (int)fi;
        const int band_k = fitem / per_band, in_band = fitem - band_k * per_band;
        const int by = in_band / P.blocks_x, bx = in_band - by * P.blocks_x;
        const int band = band_k * P.shard_count + P.shard_rank;
        const int bw = min(kBlockDim, P.w - bx * kBlockDim), bh = min(P.item_h, P.band_h - by * P.item_h);
        const int X0 = bx * kBlockDim, Y0 = P.y0 + band * P.band_h + by * P.item_h;
        if (lane == 0) samples += (unsigned)(bw * bh);                                 // every sample of the clipped block is rendered

        // x depends on the column only (Fractal2D.hs:82): the four column values of a lane are hoisted per block
        f2 xs[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int col = X0 + c * 16 + 2 * mx;
            const float q0 = KD ? div_known((float)col, P.fw, P.rcp_fw) : (float)col / P.fw;
            const float q1 = KD ? div_known((float)(col + 1), P.fw, P.rcp_fw) : (float)(col + 1) / P.fw;
            xs[c].x = q0 * 2.9f * P.ratio - P.xshift;
            xs[c].y = q1 * 2.9f * P.ratio - P.xshift;
            asm volatile("" : "+f"(xs[c].x), "+f"(xs[c].y));   // keep the values: without this ptxas re-does the IEEE divisions per sample
        }

        int y_lane = Y0 + my;                                                          // this lane's row in the first tile row
        asm volatile("" : "+r"(y_lane));             // keep it: ptxas otherwise redoes Y0 (three multiply-adds on reloaded parameters) in every tile row
#pragma unroll 1
        for (int trow = 0; trow < (P.item_h >> 2); trow++) {
            const int ty = trow * 4;
            if (ty >= bh) break;
            const int ly = ty + my;
            const float qy = KD ? div_known((float)(y_lane + ty), P.fh, P.rcp_fh) : (float)(y_lane + ty) / P.fh;
            const f2 ys = F2(qy * 2.9f - 1.45f);                                    // :81
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int lx = c * 16 + 2 * mx;
                if (c * 16 >= bw) break;                                            // warp-uniform
                unsigned live = 0u;
                if (ly < bh) live = (lx < bw ? 1u : 0u) | (lx + 1 < bw ? 2u : 0u);
                uint32_t w0 = 0u, w1 = 0u;                                           // the two pixel words (toGreen, :93-94)
                if (live) {
                    f2 jr = F2(P.julia_r), zr = xs[c], zi = ys;
                    if (live != 3u) { jr.y = kNaN; zr.y = kNaN; }                   // slot 1 alone cannot be live
                    f2 zr2 = zr * zr, zi2 = zi * zi;
                    int it0 = 40, it1 = 40;
                    float m0 = 0.0f, m1 = 0.0f;                                      // |z|² at the escape
                    unsigned done = live ^ 3u;
                    // first block: its start value is (xs, ys) itself
                    RMDF_JULIA_BLOCK(0)
                    RMDF_JULIA_PERIODIC(xs[c], ys)
                    {
                        int base = RMDF_JULIA_BLOCK_LEN;
#pragma unroll 1
                        for (; base < 40; base += RMDF_JULIA_BLOCK_LEN) {
                            const f2 sr0 = zr, si0 = zi;
                            RMDF_JULIA_BLOCK(base)
                            RMDF_JULIA_PERIODIC(sr0, si0)
                        }
                    }
                finished:
                    w0 = julia_word<SMOOTH>(it0, m0, log_2, rcp_log_2, lut_addr);
                    iters += (unsigned)it0;
                    if (live == 3u) { w1 = julia_word<SMOOTH>(it1, m1, log_2, rcp_log_2, lut_addr); iters += (unsigned)it1; }
                }
                if (live == 3u) *reinterpret_cast<uint2 *>(stage + ly * kBlockDim + lx) = make_uint2(w0, w1);
                else if (live)  stage[ly * kBlockDim + lx] = w0;
            }
        }
        __syncwarp();
        for (int l = 1; l <= k; l++) {
            const int lw = bw >> l, lh = bh >> l, h = 1 << (l - 1);
            if (bw == kBlockDim) {
                // full-width block: a lane keeps its column and walks down the rows with a constant pointer stride
                // (32 >> l outputs per row, 1 << l rows per pass): no index arithmetic inside the loop
                uint32_t *p = stage + (((lane >> (5 - l)) << l) * kBlockDim + ((lane & (lw - 1)) << l));
                const int stride = kBlockDim << (2 * l);
                for (int y = lane >> (5 - l); y < lh; y += 1 << l, p += stride)
                    p[0] = j_box4_green(p[0], p[h], p[h * kBlockDim], p[h * kBlockDim + h]);
            } else {
                for (int g = lane; g < lw * lh; g += 32) {
                    const int y = g / lw, x = g - y * lw;
                    uint32_t *p = stage + ((y << l) * kBlockDim + (x << l));
                    p[0] = j_box4_green(p[0], p[h], p[h * kBlockDim], p[h * kBlockDim + h]);
                }
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
    signal_when_grid_done(P.sync);
}
#undef RMDF_JULIA_BLOCK
#undef RMDF_JULIA_PAIR
#undef RMDF_JULIA_UPDATE
#undef RMDF_JULIA_PERIODIC

cudaError_t launch_julia(const JuliaParams &p, int grid, cudaStream_t s)
{
    if (p.known_div) {
        if (p.smooth) julia_kernel<true, true><<<grid, kCtaThreads, 0, s>>>(p);
        else          julia_kernel<false, true><<<grid, kCtaThreads, 0, s>>>(p);
    } else {
        if (p.smooth) julia_kernel<true, false><<<grid, kCtaThreads, 0, s>>>(p);
        else          julia_kernel<false, false><<<grid, kCtaThreads, 0, s>>>(p);
    }
    return cudaGetLastError();
}
int julia_ctas_per_sm()
{
    int best = 1 << 30;                        // one grid size for all variants: the smallest occupancy
    const void *fns[4] = { (const void *)julia_kernel<false, false>, (const void *)julia_kernel<true, false>,
                           (const void *)julia_kernel<false, true>, (const void *)julia_kernel<true, true> };
    for (int i = 0; i < 4; i++) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fns[i], kCtaThreads, 0) != cudaSuccess || n < 1) n = 1;
        if (n < best) best = n;
    }
    return best;
}

}  // namespace rmdf
