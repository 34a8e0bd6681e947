// envmap.cu — environment-map preprocessing kernels (compiled -fmad=false; FMAs are
// written explicitly where they are wanted).
//
//   convolve_kernel   cosineConvolveHDREnvMap, HDREnvMap.hs:217-254  (the O(n^4) hot loop)
//   resize_kernel     resizeHDRImage,          HDREnvMap.hs:169-195
//   cube_kernel       latLongHDREnvMapToCubeMap, HDREnvMap.hs:118-163 (+ CoordTransf.hs)
//   pack_half4_kernel the GL RGB16F upload,    HDREnvMap.hs:160-161
//   mip_kernel        glGenerateMipmap chain,  FrameBuffer.hs:153-154
//
// Convolution design: see the comment above convolve_kernel (persistent grid over dynamic chunks, exact skipping of
// the half of the pairs that lie behind the lobe, the reference's φ table in shared memory, powers {1,8,64,512} from
// nine squarings, packed FP32 accumulators, deterministic finish pass).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "rmdf_internal.h"
#include "packed_f32.cuh"

namespace rmdf {

constexpr int CV_DX = 32, CV_DY = 8, CV_THREADS = CV_DX * CV_DY;

// Work decomposition.  A chunk = destination tile t (32×8 pixels) × CV_ROWS source rows; chunks are ordered
// tile-major and claimed one at a time from an atomic counter by a persistent grid (as many CTAs as are resident).
// Round 1 launched 512 equal CTAs on 148 SMs (3 or 4 per SM, 86 % balance); with the skipping below the cost of a
// chunk also varies by a factor of four, so the deal has to be dynamic.  Every chunk writes one partial sum per
// destination pixel at its own place in `partial`, and the finish kernel adds a pixel's chunks in ascending source-row
// order: the result does not depend on which CTA computed what (deterministic, no float atomics).
//
// Skipping.  Half of all (dst, src) pairs have cosAngle <= 0 (the source pixel is behind the lobe) and contribute
// exactly nothing: weight 0, n unchanged.  A warp is 32 destination pixels of ONE destination row, so for a source
// row y the terms a = cosθ_L·cosθ_P and b = sinθ_L·sinθ_P are the same for all its lanes, and cosAngle = a + b·tab is
// monotonic in tab (each rounding is monotonic).  With the largest and smallest table entry of a block of 32 source
// columns × the tile's 32 destination columns (s_ext, built when the table slice is loaded) the warp knows without a
// vote — and bit-exactly — whether ANY pair of the block is positive, and skips the block if none is.
//
// srcw: the source premultiplied by sinθ_P as float4 (premultiply_kernel), read with one 128-bit load per pixel that
//       is uniform across the warp (L1-resident: a source row is 4 KB); no staging, no barrier inside a chunk
// tabT: [tiles_x][src_w][32] cos|φ_L − φ_x|, the table slice of a tile contiguous (coalesced 32 KB per chunk);
//       cth/sth: cos/sin θ per row (host, libm)
// partial: [chunks][CV_THREADS][13]  (4 powers × rgb, n)
constexpr int CV_ROWS = 4;

__global__ void premultiply_kernel(const float *__restrict__ src, int npix, int w, const float *__restrict__ sth, float4 *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float ps = sth[i / w];
    out[i] = make_float4(src[i * 3] * ps, src[i * 3 + 1] * ps, src[i * 3 + 2] * ps, 0.0f);
}

template <bool CHAIN>
__global__ void __launch_bounds__(CV_THREADS)
convolve_kernel(const float4 *__restrict__ srcw, int w, int h, const float *__restrict__ tabT,
                const float *__restrict__ cth, const float *__restrict__ sth,
                float p0, float p1, float p2, float p3, int np,
                int dst_y0, int dst_y1, int tiles_x, int chunks_per_tile, int n_chunks,
                unsigned int *__restrict__ queue, float *__restrict__ partial)
{
    extern __shared__ float smem[];
    float *s_tab = smem;                                  // [w][CV_DX]  (the tile's table slice)
    float *s_ext = smem + (size_t)w * CV_DX;              // [2][ceil(w/32)]  max / min of the slice per block of 32 source columns
    __shared__ int s_chunk;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int groups = (w + 31) >> 5;
    int loaded_tx = -1;

    for (;;) {
        __syncthreads();                                   // everybody is done with s_chunk and s_tab of the previous chunk
        if (threadIdx.x == 0) s_chunk = (int)atomicAdd(queue, 1u);
        __syncthreads();
        const int chunk = s_chunk;
        if (chunk >= n_chunks) break;
        const int tile = chunk / chunks_per_tile, y_begin = (chunk - tile * chunks_per_tile) * CV_ROWS;
        const int y_end = min(h, y_begin + CV_ROWS);
        const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
        if (tile_x != loaded_tx) {
            const float *slice = tabT + (size_t)tile_x * w * CV_DX;
            for (int i = threadIdx.x; i < w * CV_DX; i += CV_THREADS) s_tab[i] = slice[i];
            __syncthreads();
            for (int g = ty; g < groups; g += CV_DY) {     // one warp per block of 32 source columns
                float mx = -2.0f, mn = 2.0f;
                for (int sx = g * 32; sx < min(w, g * 32 + 32); sx++) { const float t = s_tab[sx * CV_DX + tx]; mx = fmaxf(mx, t); mn = fminf(mn, t); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); }
                if (tx == 0) { s_ext[g] = mx; s_ext[groups + g] = mn; }
            }
            __syncthreads();
            loaded_tx = tile_x;
        }
        // a destination row beyond dst_y1 behaves like the last valid one (warp-uniform either way), its sums are discarded
        const int dsty = min(dst_y0 + tile_y * CV_DY + ty, dst_y1 - 1);
        const float lobe_cos = cth[dsty], lobe_sin = sth[dsty];
        // accumulators: powers (1, 8) and (64, 512) side by side in 64-bit register pairs — one FFMA2 per channel and
        // pair does two individually rounded FMAs in one issue slot (csrc/packed_f32.cuh); generic powers use .x/.y too
        f2 a01[3], a23[3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) { a01[ch] = F2(0.0f); a23[ch] = F2(0.0f); }
        int n = 0;

        for (int y = y_begin; y < y_end; y++) {
            const float a = __fmul_rn(lobe_cos, cth[y]);   // thetaLobeCos * thetaPxCos      (uniform across the warp)
            const float b = __fmul_rn(lobe_sin, sth[y]);   // thetaLobeSin * thetaPxSin
            const float4 *row = srcw + (size_t)y * w;
            for (int g = 0; g < groups; g++) {
                // the block's largest cosAngle: a + b·(largest table entry if b >= 0, else the smallest)
                const float t_ext = (b >= 0.0f) ? s_ext[g] : s_ext[groups + g];
                if (!(__fadd_rn(a, __fmul_rn(b, t_ext)) > 0.0f)) continue;
                const int x_end = min(w, g * 32 + 32);
#pragma unroll 4
                for (int x = g * 32; x < x_end; x++) {
                    // cosAngle, HDREnvMap.hs:242-244, same rounding sequence as the reference
                    const float c = __fadd_rn(a, __fmul_rn(b, s_tab[x * CV_DX + tx]));
                    const float4 px = __ldg(row + x);
                    const bool pos = c > 0.0f;
                    const float cc = fmaxf(c, 0.0f);
                    n += pos ? 1 : 0;
                    f2 q01, q23;
                    if (CHAIN) {                           // powers 1, 8, 64, 512 from nine squarings
                        const float c2 = cc * cc, c4 = c2 * c2, c8 = c4 * c4;
                        const float c16 = c8 * c8, c32 = c16 * c16, c64 = c32 * c32;
                        const float c128 = c64 * c64, c256 = c128 * c128, c512 = c256 * c256;
                        q01 = F2(cc, c8); q23 = F2(c64, c512);
                    } else {
                        q01 = F2(pos ? powf(c, p0) : 0.0f, (np > 1 && pos) ? powf(c, p1) : 0.0f);
                        q23 = F2((np > 2 && pos) ? powf(c, p2) : 0.0f, (np > 3 && pos) ? powf(c, p3) : 0.0f);
                    }
                    a01[0] = fma2(px.x, q01, a01[0]); a01[1] = fma2(px.y, q01, a01[1]); a01[2] = fma2(px.z, q01, a01[2]);
                    a23[0] = fma2(px.x, q23, a23[0]); a23[1] = fma2(px.y, q23, a23[1]); a23[2] = fma2(px.z, q23, a23[2]);
                }
            }
        }
        float *o = partial + ((size_t)chunk * CV_THREADS + threadIdx.x) * 13;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) { o[ch] = a01[ch].x; o[3 + ch] = a01[ch].y; o[6 + ch] = a23[ch].x; o[9 + ch] = a23[ch].y; }
        o[12] = (float)n;
    }
}

// out[p][y][x][c] = Σ_chunks acc / Σ_chunks n   ("divide sum by number of hemisphere samples", :231-232).
// One thread per destination pixel, its tile's chunks in ascending source-row order.
__global__ void convolve_finish_kernel(const float *__restrict__ partial, int w, int npix, int np, int dst_y0, int dst_y1,
                                       int tiles_x, int chunks_per_tile, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int rows = dst_y1 - dst_y0;
    if (i >= rows * w) return;
    const int dsty = dst_y0 + i / w, dstx = i % w;
    const int tile = ((dsty - dst_y0) / CV_DY) * tiles_x + dstx / CV_DX;
    const int t = ((dsty - dst_y0) % CV_DY) * CV_DX + (dstx % CV_DX);              // thread index inside the tile's CTA
    float acc[12], n = 0.0f;
    for (int k = 0; k < 12; k++) acc[k] = 0.0f;
    for (int c = 0; c < chunks_per_tile; c++) {
        const float *p = partial + (((size_t)tile * chunks_per_tile + c) * CV_THREADS + t) * 13;
        for (int k = 0; k < 12; k++) acc[k] += p[k];
        n += p[12];
    }
    const size_t pix = (size_t)dsty * w + dstx;
    for (int p = 0; p < np; p++)
        for (int ch = 0; ch < 3; ch++) out[((size_t)p * npix + pix) * 3 + ch] = acc[p * 3 + ch] / n;
}

// Host-side tables, same libm calls and rounding as HDREnvMap.hs:223-236.  The φ table is stored per destination tile
// column block: tabT[tile_x][sx][j] = cos|φ_L(tile_x·32 + j) − φ_x(sx)|; destination columns beyond the image repeat
// the last one (their lanes compute and discard; the per-block extrema stay those of real columns).
static void convolve_tables(int w, int h, float *tabT, float *cth, float *sth)
{
    const float pi = 3.14159265358979323846f;
    for (int y = 0; y < h; y++) {
        const float theta = (float)y / (float)(h - 1) * pi;
        cth[y] = cosf(theta);
        sth[y] = sinf(theta);
    }
    const int tiles_x = (w + CV_DX - 1) / CV_DX;
    for (int t = 0; t < tiles_x; t++)
        for (int j = 0; j < CV_DX; j++) {
            const int dx = (t * CV_DX + j < w) ? t * CV_DX + j : w - 1;
            const float phi_lobe = (float)dx / (float)(w - 1) * 2.0f * pi;
            for (int sx = 0; sx < w; sx++)
                tabT[((size_t)t * w + sx) * CV_DX + j] = cosf(fabsf(phi_lobe - (float)sx / (float)(w - 1) * 2.0f * pi));
        }
}

cudaError_t launch_convolve(const float *src_dev, int w, int h, const float *powers_host, int np,
                            float *out_dev, int dst_y0, int dst_y1, cudaStream_t s)
{
    cudaError_t e;
    const int tiles_x = (w + CV_DX - 1) / CV_DX, tiles_y = (dst_y1 - dst_y0 + CV_DY - 1) / CV_DY;
    const size_t ntab = (size_t)tiles_x * w * CV_DX;
    const size_t nall = ntab + 2 * (size_t)h;
    // The tables depend on (w, h) only — w² + 2h libm calls and a 260 KB upload — so the device copy is kept per
    // (device, w, h) for the next sweep of the same size (a few sizes at most: never freed before exit).
    float *d_tab = nullptr;
    {
        struct Entry { int dev, w, h; float *d; };
        static std::mutex mu;
        static std::vector<Entry> cache;
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        std::lock_guard<std::mutex> lk(mu);
        for (const Entry &c : cache) if (c.dev == dev && c.w == w && c.h == h) d_tab = c.d;
        if (!d_tab) {
            std::vector<float> host(nall);
            convolve_tables(w, h, host.data(), host.data() + ntab, host.data() + ntab + h);
            if ((e = cudaMalloc(&d_tab, sizeof(float) * nall)) != cudaSuccess) return e;
            if ((e = cudaMemcpy(d_tab, host.data(), sizeof(float) * nall, cudaMemcpyHostToDevice)) != cudaSuccess) { cudaFree(d_tab); return e; }
            cache.push_back({ dev, w, h, d_tab });
        }
    }
    const size_t npix = (size_t)w * h;
    const int chunks_per_tile = (h + CV_ROWS - 1) / CV_ROWS;
    const int n_chunks = tiles_x * tiles_y * chunks_per_tile;
    const size_t smem = sizeof(float) * ((size_t)w * CV_DX + 2 * (size_t)((w + 31) / 32));
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(convolve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(convolve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, convolve_kernel<true>, CV_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int ctas = sms * per_sm;                               // persistent: every CTA is resident and pulls chunks
    if (ctas > n_chunks) ctas = n_chunks;
    // scratch: [64 B queue counters, one per batch of four powers][partial sums][premultiplied source]
    const int batches = (np + 3) / 4;
    const size_t q_bytes = ((size_t)batches * sizeof(unsigned int) + 255) & ~(size_t)255;
    const size_t p_bytes = sizeof(float) * (size_t)n_chunks * CV_THREADS * 13;
    unsigned char *d_scratch = nullptr;
    if ((e = cudaMallocAsync(&d_scratch, q_bytes + p_bytes + sizeof(float4) * npix, s)) != cudaSuccess) return e;
    unsigned int *d_queue = reinterpret_cast<unsigned int *>(d_scratch);
    float *d_partial = reinterpret_cast<float *>(d_scratch + q_bytes);
    float4 *d_srcw = reinterpret_cast<float4 *>(d_scratch + q_bytes + p_bytes);
    e = cudaMemsetAsync(d_queue, 0, q_bytes, s);
    if (e == cudaSuccess) {
        premultiply_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(src_dev, (int)npix, w, d_tab + ntab + h, d_srcw);
        e = cudaGetLastError();
    }
    for (int first = 0; first < np && e == cudaSuccess; first += 4) {
        const int cnt = (np - first < 4) ? np - first : 4;
        float p[4] = { 0, 0, 0, 0 };
        for (int k = 0; k < cnt; k++) p[k] = powers_host[first + k];
        const bool chain = cnt == 4 && p[0] == 1.0f && p[1] == 8.0f && p[2] == 64.0f && p[3] == 512.0f;
        if (chain)
            convolve_kernel<true><<<(unsigned)ctas, CV_THREADS, smem, s>>>(d_srcw, w, h, d_tab, d_tab + ntab, d_tab + ntab + h, p[0], p[1], p[2], p[3], cnt,
                                                                          dst_y0, dst_y1, tiles_x, chunks_per_tile, n_chunks, d_queue + first / 4, d_partial);
        else
            convolve_kernel<false><<<(unsigned)ctas, CV_THREADS, smem, s>>>(d_srcw, w, h, d_tab, d_tab + ntab, d_tab + ntab + h, p[0], p[1], p[2], p[3], cnt,
                                                                           dst_y0, dst_y1, tiles_x, chunks_per_tile, n_chunks, d_queue + first / 4, d_partial);
        e = cudaGetLastError();
        if (e != cudaSuccess) break;
        const int npx = (dst_y1 - dst_y0) * w;
        convolve_finish_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, s>>>(d_partial, w, (int)npix, cnt, dst_y0, dst_y1, tiles_x, chunks_per_tile,
                                                                             out_dev + (size_t)first * npix * 3);
        e = cudaGetLastError();
    }
    cudaFreeAsync(d_scratch, s);
    return e;
}

// ---- pixelAtBilinear, HDREnvMap.hs:91-113 ------------------------------------------------
__device__ __forceinline__ void pixel_at_bilinear(const float *img, int w, int h, float u, float v, float out[3])
{
    const float upx = u * ((float)w - 1.0f), upy = v * ((float)h - 1.0f);
    // The reference indexes with unsafePixelAt and can read one row past the image when dsth rounds up
    // (e.g. 200×103 → 100×52); a C ABI taking arbitrary sizes clamps instead — identical for every in-range sample.
    const int x = min(max((int)floorf(upx), 0), w - 1), y = min(max((int)floorf(upy), 0), h - 1);
    const int xp1 = (x + 1) % (w - 1);                       // sic: mod (w-1), :100
    const int yp1 = min(h - 1, y + 1);
    const float ur = upx - (float)x, vr = upy - (float)y, uo = 1.0f - ur, vo = 1.0f - vr;
    const float *t00 = img + ((size_t)y * w + x) * 3, *t10 = img + ((size_t)y * w + xp1) * 3;
    const float *t01 = img + ((size_t)yp1 * w + x) * 3, *t11 = img + ((size_t)yp1 * w + xp1) * 3;
#pragma unroll
    for (int c = 0; c < 3; c++)
        out[c] = (t00[c] * uo + t10[c] * ur) * vo + (t01[c] * uo + t11[c] * ur) * vr;
}

// ---- resizeHDRImage, HDREnvMap.hs:169-195 (IEEE ops only: bit-identical to the oracle) ----
__global__ void resize_kernel(const float *__restrict__ src, int srcw, int srch, int dstw, int dsth, float *__restrict__ dst)
{
    const int dstx = blockIdx.x * blockDim.x + threadIdx.x, dsty = blockIdx.y * blockDim.y + threadIdx.y;
    if (dstx >= dstw || dsty >= dsth) return;
    const float scale = (float)srcw / (float)dstw;
    const int taps = (int)ceilf(scale);
    const float ntaps = (float)(taps * taps), step = scale / (float)taps;
    const float srcx1 = (float)dstx * scale, srcy1 = (float)dsty * scale;
    float acc[3] = { 0.0f, 0.0f, 0.0f };
    for (int ty = 0; ty < taps; ty++)
        for (int tx = 0; tx < taps; tx++) {
            const float srcx = srcx1 + (float)tx * step, srcy = srcy1 + (float)ty * step;
            float p[3];
            pixel_at_bilinear(src, srcw, srch, srcx / ((float)srcw - 1.0f), srcy / ((float)srch - 1.0f), p);
            acc[0] = acc[0] + p[0]; acc[1] = acc[1] + p[1]; acc[2] = acc[2] + p[2];
        }
    float *o = dst + ((size_t)dsty * dstw + dstx) * 3;
    o[0] = acc[0] / ntaps; o[1] = acc[1] / ntaps; o[2] = acc[2] / ntaps;
}
cudaError_t launch_resize(const float *src_dev, int w, int h, int dstw, int dsth, float *out_dev, cudaStream_t s)
{
    const dim3 block(32, 8), grid((dstw + 31) / 32, (dsth + 7) / 8);
    resize_kernel<<<grid, block, 0, s>>>(src_dev, w, h, dstw, dsth, out_dev);
    return cudaGetLastError();
}

// ---- latLongHDREnvMapToCubeMap, HDREnvMap.hs:118-163 ---------------------------------------
__global__ void cube_kernel(const float *__restrict__ latlong, int w, int h, int fw, float *__restrict__ faces)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, face = blockIdx.z;
    if (x >= fw || y >= fw) return;
    const float pi = 3.14159265358979323846f;
    // cubeMapPixelToDir, :76-87
    const float vw = ((float)x + 0.5f) / (float)fw * 2.0f - 1.0f;
    const float vh = ((float)y + 0.5f) / (float)fw * 2.0f - 1.0f;
    float d0, d1, d2;
    switch (face) {
    case 0:  d0 = 1.0f;  d1 = -vh;   d2 = -vw;   break;
    case 1:  d0 = -1.0f; d1 = -vh;   d2 = vw;    break;
    case 2:  d0 = vw;    d1 = 1.0f;  d2 = vh;    break;
    case 3:  d0 = vw;    d1 = -1.0f; d2 = -vh;   break;
    case 4:  d0 = vw;    d1 = -vh;   d2 = 1.0f;  break;
    default: d0 = -vw;   d1 = -vh;   d2 = -1.0f; break;
    }
    // linear's normalize: unchanged when |l| or |1-l| <= 1e-6
    const float l = (d0 * d0 + d1 * d1) + d2 * d2;
    if (!(fabsf(l) <= 1e-6f || fabsf(1.0f - l) <= 1e-6f)) {
        const float s = sqrtf(l);
        d0 = d0 / s; d1 = d1 / s; d2 = d2 / s;
    }
    // worldToLocal (CoordTransf.hs:47-51): (x, -z, y); cartesianToSpherical (:35-45)
    const float lx = d0, ly = -d2, lz = d1;
    const float theta = acosf(fmaxf(-1.0f, fminf(1.0f, lz)));
    float phi = atan2f(ly, lx);
    if (phi < 0.0f) phi = phi + 2.0f * pi;
    if (phi == 2.0f * pi) phi = 0.0f;
    // sphericalToEnvironmentUV (:61-71)
    const float p1 = phi + pi / 2.0f;
    const float p2 = (p1 > 2.0f * pi) ? p1 - 2.0f * pi : p1;
    const float p3 = 2.0f * pi - p2;
    float col[3];
    pixel_at_bilinear(latlong, w, h, p3 / (pi * 2.0f), theta / pi, col);
    float *o = faces + (((size_t)face * fw + y) * fw + x) * 3;
    // GL converts to half on upload (RGB16F, :160-161); keep the rounded value as float
    o[0] = __half2float(__float2half_rn(col[0]));
    o[1] = __half2float(__float2half_rn(col[1]));
    o[2] = __half2float(__float2half_rn(col[2]));
}
cudaError_t launch_latlong_to_cube(const float *latlong_dev, int w, int h, int face_w, float *faces_dev, cudaStream_t s)
{
    const dim3 block(32, 8), grid((face_w + 31) / 32, (face_w + 7) / 8, 6);
    cube_kernel<<<grid, block, 0, s>>>(latlong_dev, w, h, face_w, faces_dev);
    return cudaGetLastError();
}

// float RGB faces → half4 texels (CUDA has no 3-channel half format), written straight into the slot's two arrays (layered
// for the parity path's exact fetches, cube map for the fast path's hardware filtering) through their surface objects;
// also rounds the float faces to half in place (rmdf_envmap_get_cube returns what the textures hold)
__global__ void pack_half4_kernel(float *__restrict__ faces, int fw, cudaSurfaceObject_t layered, cudaSurfaceObject_t cube)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, face = blockIdx.z;
    if (x >= fw) return;
    const size_t i = ((size_t)face * fw + y) * fw + x;
    const __half r = __float2half_rn(faces[i * 3]), g = __float2half_rn(faces[i * 3 + 1]), b = __float2half_rn(faces[i * 3 + 2]);
    faces[i * 3] = __half2float(r); faces[i * 3 + 1] = __half2float(g); faces[i * 3 + 2] = __half2float(b);
    const ushort4 t = make_ushort4(__half_as_ushort(r), __half_as_ushort(g), __half_as_ushort(b), 0x3c00 /* 1.0 */);
    surf2DLayeredwrite(t, layered, x * (int)sizeof(ushort4), y, face);
    surf2DLayeredwrite(t, cube, x * (int)sizeof(ushort4), y, face);       // a cube map array is a 6-layer surface, layer = face
}
cudaError_t launch_pack_half4(const float *faces_dev, int face_w, cudaSurfaceObject_t layered, cudaSurfaceObject_t cube, cudaStream_t s)
{
    const dim3 grid((unsigned)(face_w + 127) / 128, (unsigned)face_w, 6);
    pack_half4_kernel<<<grid, 128, 0, s>>>(const_cast<float *>(faces_dev), face_w, layered, cube);
    return cudaGetLastError();
}

// ---- glGenerateMipmap box chain, one level per launch ---------------------------------------
__device__ __forceinline__ uint32_t mip_box4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    uint32_t lo = (a & 0x00ff00ffu) + (b & 0x00ff00ffu) + (c & 0x00ff00ffu) + (d & 0x00ff00ffu) + 0x00020002u;
    uint32_t hi = ((a >> 8) & 0x00ff00ffu) + ((b >> 8) & 0x00ff00ffu) + ((c >> 8) & 0x00ff00ffu) + ((d >> 8) & 0x00ff00ffu) + 0x00020002u;
    return ((lo >> 2) & 0x00ff00ffu) | (((hi >> 2) & 0x00ff00ffu) << 8);
}
__global__ void mip_kernel(const uint32_t *__restrict__ in, int w, int nw, int nh, uint32_t *__restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nw || y >= nh) return;
    const uint32_t *p = in + (size_t)(2 * y) * w + 2 * x;
    out[(size_t)y * nw + x] = mip_box4(p[0], p[1], p[w], p[w + 1]);
}
cudaError_t launch_mip_resolve(const uint32_t *in_dev, int w, int h, int levels, uint32_t *out_dev, cudaStream_t s)
{
    if (levels == 0) return cudaMemcpyAsync(out_dev, in_dev, sizeof(uint32_t) * (size_t)w * h, cudaMemcpyDeviceToDevice, s);
    uint32_t *tmp[2] = { nullptr, nullptr };
    cudaError_t e = cudaSuccess;
    const uint32_t *cur = in_dev;
    for (int l = 0; l < levels && e == cudaSuccess; l++) {
        const int nw = w >> 1, nh = h >> 1;
        uint32_t *dst = out_dev;
        if (l != levels - 1) {
            if (!tmp[l & 1]) e = cudaMallocAsync(&tmp[l & 1], sizeof(uint32_t) * (size_t)nw * nh, s);
            if (e != cudaSuccess) break;
            dst = tmp[l & 1];
        }
        const dim3 block(32, 8), grid((nw + 31) / 32, (nh + 7) / 8);
        mip_kernel<<<grid, block, 0, s>>>(cur, w, nw, nh, dst);
        e = cudaGetLastError();
        cur = dst; w = nw; h = nh;
    }
    if (tmp[0]) cudaFreeAsync(tmp[0], s);
    if (tmp[1]) cudaFreeAsync(tmp[1], s);
    return e;
}

}  // namespace rmdf
