/* rmdf.h — C ABI of librmdf_b200.so, the B200 (sm_100a) replacement for the one hot
 * path of blitzcode/ray-marching-distance-fields: per-pixel sphere tracing
 * (fragment.shd), the environment-map prefilter (HDREnvMap.hs) and the 2D Julia set
 * (Fractal2D.hs).
 *
 * The reference has no FFI layer today; its seams are three Haskell
 * signatures.  Every entry point below names the reference interface it replaces
 * (file:line under /root/reference) and is what a `foreign import ccall` in
 * ShaderRendering.hs / FrameBuffer.hs / HDREnvMap.hs binds (INTEGRATION.md shows the
 * stubs; haskell/RMDF/FFI.hs holds them).
 *
 * Conventions
 *   - Plain pointers and sizes only.  The caller owns every buffer it passes; the
 *     library never frees or retains caller memory after a call returns.
 *   - Every call returns 0 on success, non-zero on error; rmdf_last_error() then
 *     returns a thread-local message (the Haskell shim turns it into `Left String`,
 *     as ShaderRendering.hs:110 does).  No exceptions cross the boundary.
 *   - There is no CPU fallback: every compute entry point fails with
 *     RMDF_ERR_CUDA when no sm_100 device is usable.
 *   - Calls may arrive on any OS thread (`safe` foreign calls); each entry point
 *     selects the context's device and serialises on the context's stream.
 *   - Images: env maps are interleaved RGB float32, row 0 = top (JuicyPixels
 *     `Image PixelRGBF`, HDREnvMap.hs:238-239).  Frame buffers are 32-bit words
 *     0xAABBGGRR, index px + py·w, row 0 = BOTTOM (the GL texture layout that
 *     fillFrameBuffer hands out, FrameBuffer.hs:117-158).
 */
#ifndef RMDF_H
#define RMDF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMDF_ABI_VERSION 1

typedef struct rmdf_ctx rmdf_ctx;

enum {
    RMDF_OK = 0,
    RMDF_ERR_ARG = 1,        /* invalid argument (message says which) */
    RMDF_ERR_CUDA = 2,       /* CUDA runtime error / no usable device */
    RMDF_ERR_STATE = 3,      /* e.g. Cornell scene without geometry */
    RMDF_ERR_UNSUPPORTED = 4
};

/* `FragmentShader` enum order, ShaderRendering.hs:46-47; shader variants :119-123 */
enum {
    RMDF_SCENE_CORNELL_BOX = 0,  /* FSDECornellBoxShader  (#define CORNELL_BOX_SCENE)       */
    RMDF_SCENE_DE_TEST = 1,      /* FSDETestShader        (no define: CSG primitive scene)  */
    RMDF_SCENE_MB_POWER8 = 2,    /* FSMBPower8Shader      (MANDELBULB_SCENE + POWER8)       */
    RMDF_SCENE_MB_GENERAL = 3    /* FSMBGeneralShader     (MANDELBULB_SCENE)                */
};

/* texture-unit order of srEnvCubeMaps, ShaderRendering.hs:86-91,173-174 */
enum {
    RMDF_ENV_REFLECTION = 0, RMDF_ENV_COS_1 = 1, RMDF_ENV_COS_8 = 2,
    RMDF_ENV_COS_64 = 3, RMDF_ENV_COS_512 = 4, RMDF_ENV_SLOTS = 5
};

enum { RMDF_MODE_PARITY = 0,   /* IEEE-only arithmetic, bit-identical to the CPU oracle */
       RMDF_MODE_FAST = 1 };   /* FMA contraction + MUFU approximations + HW filtering  */
enum { RMDF_OUT_RGBA8 = 0,     /* gamma-encoded 0xAABBGGRR, what the GL FBO holds       */
       RMDF_OUT_RGBA32F = 1 }; /* linear pre-gamma RGB + hit flag in A (test/inspection) */
enum { RMDF_LOD_QUAD = 0,      /* min=NEAREST/mag=LINEAR per 2×2 quad (GLHelpers.hs:105) */
       RMDF_LOD_ALWAYS_LINEAR = 1 };
/* Switches that exist in fragment.shd but are OFF in the reference's build.
 * Each is bit-identical to the oracle with the same switch; none changes the default path. */
enum { RMDF_FLAG_NORMAL_CENTRAL = 1,    /* normal_central_difference (fragment.shd:471-480) instead of :463-470 */
       RMDF_FLAG_CAM_ORTHO = 2,         /* #define CAM_ORTHO (fragment.shd:906-908, 853-861), width 2 */
       RMDF_FLAG_RAY_TRACING_TEST = 4, /* #define RAY_TRACING_TEST (fragment.shd:915-937): ray-traced Cornell box */
       /* NOT IN THE REFERENCE (fragment.shd has no shadow code): BASELINE.json's north_star names soft
        * shadows, so they exist as an opt-in that is off by default and excluded from every parity statement.  A
        * distance-field penumbra march (Quilez: res = min(res, k·d/t), k = 8, ≤ 48 steps) from the hit point towards
        * the fixed light normalize(1,1,1) of the shader's commented-out soft_lam call (:778); scales the diffuse and
        * specular terms of :805-810 by the result.  Hit/miss masks and miss pixels are unchanged. */
       RMDF_FLAG_SOFT_SHADOWS = 8 };
enum { RMDF_LAYOUT_FRAME = 0,  /* output indexed px + py·w over the whole frame          */
       RMDF_LAYOUT_PACKED_TILES = 1 }; /* the rendered tiles back to back, ascending (ty, tx) */

/* ---- lifetime: withShaderRenderer, ShaderRendering.hs:60-110 ------------------- */
/* One context per GPU.  `device` is a CUDA ordinal.  Owns a stream, the five env-map
 * texture objects, the Cornell table, the work-queue counter and the counters. */
int  rmdf_create(rmdf_ctx **out, int device);
void rmdf_destroy(rmdf_ctx *ctx);
/* `Either String`: message of the last failing call on this thread ("" if none). */
const char *rmdf_last_error(void);
/* The same for the last failing call ON THIS CONTEXT, copied into buf (NUL-terminated, at most cap bytes), whatever
 * OS thread made that call: under GHC -threaded an unbound Haskell thread may run two consecutive `safe` foreign
 * calls on different OS threads, so the shim (haskell/RMDF/FFI.hs `check`) fetches the message through the context. */
int  rmdf_ctx_last_error(rmdf_ctx *ctx, char *buf, int cap);
int  rmdf_abi_version(void);
/* Run on the caller's CUDA stream (cudaStream_t) instead of the context's own;
 * NULL restores the private stream.  Lets a host time with its own events. */
int  rmdf_set_stream(rmdf_ctx *ctx, void *cuda_stream);
int  rmdf_sync(rmdf_ctx *ctx);

/* ---- env-map preprocessing: HDREnvMap.hs --------------------------------------- */
/* resizeHDRImage (HDREnvMap.hs:169-195).  out_h = round(h/w·dstw); out holds
 * dstw·out_h·3 floats.  rmdf_envmap_resize_height returns out_h. */
int  rmdf_envmap_resize_height(int w, int h, int dstw);
int  rmdf_envmap_resize(rmdf_ctx *ctx, const float *rgb, int w, int h, int dstw, float *out);
/* cosineConvolveHDREnvMap (HDREnvMap.hs:217-254) for `np` powers in ONE pass
 * (replaces the mapConcurrently over powers, ShaderRendering.hs:142-144).
 * out holds np·w·h·3 floats, power-major.  Re-entrant. */
int  rmdf_envmap_convolve(rmdf_ctx *ctx, const float *rgb, int w, int h,
                          const float *powers, int np, float *out);
/* Same, with source and destination already in device memory on ctx's device. */
int  rmdf_envmap_convolve_device(rmdf_ctx *ctx, const void *rgb_dev, int w, int h,
                                 const float *powers, int np, void *out_dev);
/* Multi-GPU decomposition of the sweep: only destination rows [row_begin, row_end) are
 * computed and written (in place inside the full np·w·h·3 output); every rank reads the whole source. */
int  rmdf_envmap_convolve_rows_device(rmdf_ctx *ctx, const void *rgb_dev, int w, int h,
                                      const float *powers, int np, int row_begin, int row_end, void *out_dev);
/* The RGBE quantisation a cache file applies (JP.saveRadianceImage then loadHDRImage,
 * ShaderRendering.hs:84,147): in-place on npix RGB triples (host memory). */
int  rmdf_rgbe_roundtrip(float *rgb, int64_t npix);
/* latLongHDREnvMapToCubeMap + setTextureShader (HDREnvMap.hs:118-163,
 * ShaderRendering.hs:81-91,173-174): lat-long → 6 faces of width w/3, RGB16F, bound
 * to texture unit `slot`. */
int  rmdf_envmap_bind(rmdf_ctx *ctx, int slot, const float *latlong_rgb, int w, int h);
/* Bind ready-made faces (6·face_w·face_w·3 floats, order +X,-X,+Y,-Y,+Z,-Z, row 0 =
 * t 0); values are rounded to half like GL's RGB16F upload. */
int  rmdf_envmap_bind_cube(rmdf_ctx *ctx, int slot, const float *faces_rgb, int face_w);
/* The same for n slots in one batch with ONE synchronisation at the end (per-frame uploads of all five maps). */
int  rmdf_envmap_bind_cubes(rmdf_ctx *ctx, int n, const int *slots, const float *const *faces_rgb, const int *face_w);
/* Read a bound cube map back (the half values widened to float). faces_rgb may be
 * NULL to query face_w only. */
int  rmdf_envmap_get_cube(rmdf_ctx *ctx, int slot, float *faces_rgb, int *face_w);
/* Unbound slots sample as constant 1.0 ("no env lighting"). */
int  rmdf_envmap_unbind(rmdf_ctx *ctx, int slot);

/* ---- Cornell geometry: mkCornellBoxVerticesTex, CornellBox.hs:21-46 ------------ */
/* 96 vertices (32 triangles) xyz float32 = the cornell_geom RGB32F texels. */
int  rmdf_set_cornell(rmdf_ctx *ctx, const float *xyz, int nverts);
/* The reference's own table, already scaled ((v/279.6 − 1)·0.99/√3), into xyz[288]. */
int  rmdf_cornell_box_vertices(float *xyz);

/* ---- frame rendering: drawShaderTile, ShaderRendering.hs:151-196 --------------- */
typedef struct {
    int32_t scene;          /* RMDF_SCENE_*  (shdEnum) */
    int32_t fb_w, fb_h;     /* frame-buffer size = in_screen_wdh/hgt; already includes the
                               super-sampling scale (App.hs:124-134) */
    float   time;           /* in_time (realToFrac of the Double, :171) */
    int32_t tile_idx;       /* Maybe Int: -1 = Nothing (full frame), else idx mod 64
                               selects tile (idx mod 8, idx div 8) of the 8×8 grid, row 0
                               at the bottom (:183-193) */
    int32_t resolve_log2;   /* 0: write FB level 0 (one word per sample, the reference
                               behaviour).  k>0: write mip level k instead — the 2×2 box
                               chain glGenerateMipmap builds (FrameBuffer.hs:153-154,
                               187-195) folded into the kernel; output is (fb_w>>k)×(fb_h>>k) */
    int32_t out_format;     /* RMDF_OUT_*  (RGBA32F requires resolve_log2 = 0) */
    int32_t mode;           /* RMDF_MODE_* */
    int32_t lod_policy;     /* RMDF_LOD_* */
    int32_t shard_rank;     /* multi-GPU: render only the tiles (tx, ty) of the shard grid with */
    int32_t shard_count;    /*   (tx + shard_skew·ty) mod shard_count == shard_rank; count <= 1 = all */
    int32_t layout;         /* RMDF_LAYOUT_* */
    int32_t shard_grid_x;   /* shard grid; 0 = the reference's 8×8 tile grid with skew 1.  A finer   */
    int32_t shard_grid_y;   /*   grid (e.g. 32×16, skew 2) balances the ranks better: the Mandelbulb's */
    int32_t shard_skew;     /*   cost is concentrated mid-frame.  shard_grid_x % shard_count == 0.    */
    int32_t flags;          /* RMDF_FLAG_*; 0 = the shader exactly as the reference builds it */
    int32_t row_begin;      /* render only frame-buffer rows [row_begin, row_end) of the selected work   */
    int32_t row_end;        /*   (0,0 = all).  Lets a host pipeline a frame in horizontal stripes, e.g. copy */
                            /*   stripe i to the CPU while stripe i+1 is marched (measured on config 4: four  */
                            /*   stripes cost as much in launch tails as the overlap saves).  Multiples of 2<<resolve_log2 */
                            /*   (of the shard tile height when sharded); RMDF_LAYOUT_FRAME only.              */
} rmdf_params;

/* Number of 32-bit words (RGBA8) or float4 (RGBA32F) the call will write, and the
 * number of tiles it renders; for sizing the caller's buffer. */
int64_t rmdf_render_output_pixels(const rmdf_params *p, int32_t *n_tiles);
/* Render into host memory (the MVector Word32 of fillFrameBuffer, FrameBuffer.hs:117-
 * 135); returns when `fb` is complete.  Pageable memory goes through a device staging
 * buffer and one device→host copy; page-locked memory (cudaHostAlloc / cudaHostRegister,
 * like the mapped PBO the reference hands out) is written directly by the kernel. */
int  rmdf_render(rmdf_ctx *ctx, const rmdf_params *p, void *fb_host);
/* Render into device memory on ctx's device; asynchronous on the context's stream. */
int  rmdf_render_device(rmdf_ctx *ctx, const rmdf_params *p, void *fb_dev);

/* ---- 2D Julia set: juliaAnimated, Fractal2D.hs:63-98 --------------------------- */
/* Rows [y0,y1) of the w×h frame, green-only words 0x0000GG00; resolve_log2 as above
 * (w, h, y0, y1 are level-0 sizes and must be multiples of 1<<resolve_log2). */
int  rmdf_julia(rmdf_ctx *ctx, int w, int h, int smooth, double tick,
                int y0, int y1, int resolve_log2, uint32_t *fb_host);
int  rmdf_julia_device(rmdf_ctx *ctx, int w, int h, int smooth, double tick,
                       int y0, int y1, int resolve_log2, void *fb_dev);
/* Multi-GPU decomposition of the same frame (the reference splits rows with
 * forSegmentsConcurrently, Fractal2D.hs:98): the h rows are cut into bands of band_h rows and
 * this call renders the bands b with b mod shard_count == shard_rank (interleaving balances the
 * ranks: interior rows cost 40 iterations, exterior few).  layout = RMDF_LAYOUT_PACKED_TILES:
 * this rank's bands back to back in ascending b; RMDF_LAYOUT_FRAME: each band at its place in the
 * whole (resolved) frame, e.g. a shared frame buffer on another GPU (rmdf_shared_open).
 * h/band_h must be a multiple of shard_count, band_h of 1<<resolve_log2. */
int  rmdf_julia_bands_device(rmdf_ctx *ctx, int w, int h, int smooth, double tick, int resolve_log2,
                             int band_h, int shard_rank, int shard_count, int layout, void *fb_dev);

/* ---- frame-buffer helpers: FrameBuffer.hs -------------------------------------- */
/* glGenerateMipmap chain (FrameBuffer.hs:153-154): level `levels` of a w×h RGBA8
 * image already on the device. */
int  rmdf_mip_resolve_device(rmdf_ctx *ctx, const void *fb_dev, int w, int h, int levels, void *out_dev);

/* ---- multi-GPU: one frame buffer that every rank's kernel writes into ------------ */
/* The gather of a sharded frame fused into the march kernel's store phase: rank 0 owns the
 * (resolved) frame buffer, the other ranks map it and pass it as `fb_dev` with
 * RMDF_LAYOUT_FRAME + shard_* — their 128-bit tile stores then travel over NVLink while the
 * rest of the frame is still being marched; only a barrier remains afterwards.  (No
 * counterpart in the reference, which is single-GPU.)
 * rmdf_shared_alloc: cudaMalloc on ctx's device + an IPC handle (64 opaque bytes) to send to
 * the other processes; rmdf_shared_open: map such a handle on ctx's device (peer access over
 * NVLink/PCIe); rmdf_enable_peer: the same-process equivalent for a second context. */
int  rmdf_shared_alloc(rmdf_ctx *ctx, int64_t bytes, void **dev_ptr, unsigned char handle[64]);
int  rmdf_shared_free(rmdf_ctx *ctx, void *dev_ptr);
int  rmdf_shared_open(rmdf_ctx *ctx, const unsigned char handle[64], void **dev_ptr);
int  rmdf_shared_close(rmdf_ctx *ctx, void *dev_ptr);
int  rmdf_enable_peer(rmdf_ctx *ctx, int peer_device);
/* A frame buffer in HOST memory that several processes' kernels write into: page-lock `bytes` at `host_ptr`
 * (e.g. a POSIX shared-memory mapping that every rank has opened) for ctx's device and return the address the
 * device uses for it; pass that as `fb_dev` with RMDF_LAYOUT_FRAME + shard_*.  Each GPU then stores its tiles
 * over its own PCIe link straight into the final host frame: no gather and no device→host copy on rank 0. */
int  rmdf_host_register(rmdf_ctx *ctx, void *host_ptr, int64_t bytes, void **dev_ptr);
int  rmdf_host_unregister(rmdf_ctx *ctx, void *host_ptr);

/* ---- counters (new; the reference only prints FPS, App.hs:221-238) ------------- */
typedef struct {
    uint64_t samples;       /* rays generated                                   */
    uint64_t in_sphere;     /* rays that entered the bounding sphere            */
    uint64_t hits;          /* rays that hit                                    */
    uint64_t march_de;      /* distance-estimator evaluations in ray_march      */
    uint64_t shade_de;      /* ... in normal_backward_difference + distance_ao  */
    uint64_t inner_iters;   /* Mandelbulb iterations / Julia iterations         */
    uint64_t launches;      /* kernels launched by this context so far          */
    float    last_kernel_ms;/* device time of the last render/julia/convolve kernel */
    float    reserved;
} rmdf_counters;
/* Totals since the last rmdf_counters_reset (synchronises the stream). */
int  rmdf_counters_get(rmdf_ctx *ctx, rmdf_counters *out);
int  rmdf_counters_reset(rmdf_ctx *ctx);

/* Work order (no reference counterpart; scheduling only, never changes a pixel).  A renderer draws frames in sequence:
 * the render kernel times its work items, one small kernel after it sorts them by the octave of their cost, and the next
 * frame with the same work list (same frame size, tile or shard, scene) claims them most-expensive-first — the frame
 * ends on its cheapest items instead of idling behind the last crevice of the fractal, and the heaviest items are
 * spread evenly over the SMs.  Measured: config 4 -0.35 %, one GPU's eighth of it -2.5 %, the general-power scene
 * -5 %, config 1 -1..4 %.  On by default for the two Mandelbulb scenes and frames of >= 4096 work items (the Cornell and
 * CSG scenes gain nothing); 0 switches it off (items in scan order, as a first frame always is). */
int  rmdf_set_work_order(rmdf_ctx *ctx, int enabled);

/* ---- one frame rendered by several GPUs, assembled over NVLink peer memory (no reference counterpart) ----------
 * A control block of RMDF_SYNC_WORDS uint32 words (zero-initialised) lives in memory that every participating GPU can
 * reach — allocate it on rank 0 with rmdf_shared_alloc next to the frame buffer, map it elsewhere with
 * rmdf_shared_open.  Word 0 = work queue, word 1 = "go", words 8+r = arrival flag of rank r.  Per frame, with a
 * sequence number seq that grows by one (start at 1):
 *   every rank   rmdf_frame_begin(ctx, ctrl, rank, world, seq, deal, timeout_ms)   rank 0's stream zeroes the queue and
 *                                                            publishes seq ("go"); arms this context's next
 *                                                            rmdf_render_device / rmdf_julia_device call:
 *   every rank   rmdf_render_device(ctx, params, frame)      the SAME full-frame params on every rank (RMDF_LAYOUT_FRAME,
 *                                                            shard_count 1), `frame` = the shared buffer.  Ranks != 0 first
 *                                                            wait on their stream for "go".  RMDF_DEAL_INTERLEAVED: work
 *                                                            item i (a 32×8-sample block) is rendered by rank i mod world
 *                                                            through the rank's own queue — no remote atomics.
 *                                                            RMDF_DEAL_QUEUE: all GPUs claim items from the one queue word
 *                                                            with system-scope atomics over NVLink.  Either way the pixels
 *                                                            are stored into the shared frame and the last CTA of each
 *                                                            kernel publishes seq in the rank's arrival flag.
 *   rank 0       rmdf_frame_end(ctx, ctrl, world, seq, timeout_ms)   its stream waits for all arrival flags: work queued
 *                                                            behind it sees the complete frame.  No host round trip.
 * Waits give up after timeout_ms (default 10 s when <= 0); rmdf_frame_sync_status synchronises the stream and returns
 * 0, or 1 + the index of the word a wait gave up on. */
#define RMDF_SYNC_WORDS 64
enum { RMDF_DEAL_INTERLEAVED = 0, RMDF_DEAL_QUEUE = 1 };
int  rmdf_frame_begin(rmdf_ctx *ctx, void *ctrl_dev, int rank, int world, uint32_t seq, int deal, int timeout_ms);
int  rmdf_frame_end(rmdf_ctx *ctx, void *ctrl_dev, int world, uint32_t seq, int timeout_ms);
int  rmdf_frame_sync_status(rmdf_ctx *ctx, int *timed_out_word);

/* FP32 FMA roofline denominator: an FFMA-chain microbenchmark on ctx's device (the
 * reference has no counterpart).  tflops = 2·FFMA/s, sm_mhz = SM clock sustained. */
int  rmdf_measure_fp32_peak(rmdf_ctx *ctx, float *tflops, float *sm_mhz);
/* MUFU (XU pipe) roofline denominator for the transcendental-bound scenes: ex2.approx chains;
 * gops = 1e9 MUFU lane-operations per second (16 lanes/clk/SM on sm_100a). */
int  rmdf_measure_mufu_peak(rmdf_ctx *ctx, float *gops);

#ifdef __cplusplus
}
#endif
#endif /* RMDF_H */
