// rmdf_internal.h — types shared by the C-ABI layer (rmdf_api.cu) and the kernel
// translation units.  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "frame_sync.cuh"

namespace rmdf {

constexpr int kBlockDim     = 32;    // a work item is a 32×32 block of frame-buffer samples
#ifndef RMDF_WARPS_PER_CTA
#define RMDF_WARPS_PER_CTA 4   // small CTAs: register/shared-memory occupancy in steps of 4 warps
#endif
constexpr int kWarpsPerCta  = RMDF_WARPS_PER_CTA;
constexpr int kCtaThreads   = kWarpsPerCta * 32;
constexpr int kCornellStride = 8;     // float4 per triangle: 6 of precomputed terms + bounding sphere + plane
constexpr int kCornellPairs = 16;                      // triangles 2k, 2k+1 are one quad of CornellBox.hs: one plane test for both
constexpr int kCornellF4    = 32 * kCornellStride + kCornellPairs;   // + one supporting plane per pair

enum CounterSlot { C_SAMPLES = 0, C_IN_SPHERE, C_HITS, C_MARCH_DE, C_SHADE_DE, C_ITERS, C_NSLOTS = 8 };

// Everything one frame needs, passed by value as the kernel parameter.
struct MarchParams {
    int   scene, fb_w, fb_h;
    float fw, fh, aspect, fov_xs;     // in_screen_wdh/hgt as float, w/h, tan(hfov/2)
    float cam[12];                    // x̂, ŷ, ẑ, eye  (lookat, fragment.shd:829-838)
    float power;                      // animated Mandelbulb power (fragment.shd:116-119)
    float bsphere_r;                  // fragment.shd:640-649
    float opaque_zero;                // always 0.0f: a zero the compiler cannot fold or re-materialise (fast power-8 loop)
    int   lod_policy, resolve_log2, out_format, layout;
    int   flags;                      // RMDF_FLAG_* (0 = the reference's build)
    // work list: the whole frame (0), one tile of the reference's 8×8 grid (1), or this rank's
    // tiles of the shard grid (2): (tx + skew·ty) mod count == rank, packed by ascending (ty, tx)
    int   rect_mode;
    int   tile_x0, tile_y0;           // rect_mode 1: origin of the rendered rect (the tile grown to even pixel boundaries)
    int   clip_x0, clip_y0, clip_x1, clip_y1;   // rect_mode 1: the pixels that are stored (the tile itself)
    int   clipped;                    // rect_mode 1: the rendered rect is larger than the stored one (odd tile edges)
    long long rect_out_base;          // rect_mode 1: output index of the rendered rect's origin
    int   shard_gx, shard_count, shard_rank, shard_skew, shard_ty0;   // shard_ty0: first tile row of a stripe
    int   tile_w, tile_h;             // level-0 size of one tile
    int   item_h;                     // rows per work item: 4, 8, 16 or 32
    int   item_w;                     // columns per work item: 32, or 16 / 8 on frames too small to fill the machine
    int   blocks_x, blocks_y;         // 32×item_h blocks per rect
    unsigned long long blocks_x_inv;  // floor(2^40 / blocks_x) + 1 when that divides every in-rect index exactly, else 0 (decode_item)
    int   n_items;
    int   out_stride;                 // output row stride in pixels (already >> resolve_log2)
    // resources
    cudaTextureObject_t env_point[5]; // layered 2D half4, point sampled (parity path)
    cudaTextureObject_t env_cube[5];  // cube map half4, linear + seamless (fast path)
    cudaTextureObject_t env_cube_nearest[5];  // cube map half4, point sampled (fast path, minified quads)
    int   env_face[5];                // face width, 0 = unbound (samples as constant 1.0)
    const float4 *cornell;            // kCornellF4 float4 (precomputed triangle terms)
    void *out;
    FrameSync sync;                   // work queue + completion signal (frame_sync.cuh)
    unsigned long long *counters;     // C_NSLOTS totals
};

struct JuliaParams {
    int   w, h, y0, y1, smooth, resolve_log2;
    int   frame_layout;                       // 1: bands land at their frame position (shared frame buffer)
    int   band_h, shard_rank, shard_count;   // rows [y0,y1) cut into bands of band_h rows; this rank renders bands b ≡ rank (mod count)
    float julia_r, julia_i, fw, fh, ratio, xshift;
    float rcp_fw, rcp_fh;                     // RN(1/fw), RN(1/fh) for the known-divisor divisions
    int   known_div;                          // neither fw's nor fh's significand is all ones: div_known is exact
    int   item_h, blocks_x, blocks_y, n_items;
    int   out_stride;
    uint32_t *out;
    FrameSync sync;
    unsigned long long *counters;
};

// launchers implemented in march_parity.cu / march_fast.cu / julia.cu / envmap.cu
cudaError_t launch_march_parity(const MarchParams &p, int grid, cudaStream_t s);
cudaError_t launch_march_fast(const MarchParams &p, int grid, cudaStream_t s);
int         march_parity_ctas_per_sm(int scene, int flags);
int         march_fast_ctas_per_sm(int scene, int flags);
cudaError_t launch_julia(const JuliaParams &p, int grid, cudaStream_t s);
int         julia_ctas_per_sm();
cudaError_t launch_convolve(const float *src_dev, int w, int h, const float *powers_host, int np,
                            float *out_dev, int dst_y0, int dst_y1, cudaStream_t s);
cudaError_t launch_resize(const float *src_dev, int w, int h, int dstw, int dsth, float *out_dev, cudaStream_t s);
cudaError_t launch_latlong_to_cube(const float *latlong_dev, int w, int h, int face_w, float *faces_dev, cudaStream_t s);
cudaError_t launch_pack_half4(const float *faces_dev, int face_w, cudaSurfaceObject_t layered, cudaSurfaceObject_t cube, cudaStream_t s);
cudaError_t measure_fp32_peak(int sm_count, cudaStream_t s, float *tflops, float *sm_mhz);
cudaError_t measure_mufu_peak(int sm_count, cudaStream_t s, float *gops);
cudaError_t preload_frame_sync_kernels();
cudaError_t launch_frame_begin(unsigned int *ctrl, unsigned int seq, cudaStream_t s);
cudaError_t launch_wait_flags(const unsigned int *flags, int n, unsigned int value, unsigned int timeout_ms, unsigned int *status, cudaStream_t s);
cudaError_t launch_order_place(const unsigned int *rec, unsigned int n_items, unsigned int *sync, unsigned int *order, cudaStream_t s);
cudaError_t launch_mip_resolve(const uint32_t *in_dev, int w, int h, int levels, uint32_t *out_dev, cudaStream_t s);

}  // namespace rmdf
