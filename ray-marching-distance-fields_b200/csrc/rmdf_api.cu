// rmdf_api.cu — the C ABI of librmdf_b200.so (include/rmdf.h): context lifetime,
// env-map texture objects, parameter validation, work-list construction, launches.
// Host arithmetic that must agree bit-for-bit with the kernels' inputs (camera,
// Cornell triangle terms) uses include/rmdf_glsl.h and is compiled with
// -Xcompiler -ffp-contract=off.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include <mutex>
#include <vector>
#include "../../include/rmdf.h"
#include "../../include/rmdf_glsl.h"
#include "rmdf_internal.h"

using namespace rmdf;

// ---------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local bool g_err_fresh = false;       // set by fail(); a DeviceGuard copies the message into its context

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    g_err_fresh = true;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(RMDF_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

struct EnvSlot {
    int face_w = 0;
    cudaArray_t arr_layered = nullptr, arr_cube = nullptr;
    cudaTextureObject_t tex_point = 0, tex_cube = 0, tex_cube_nearest = 0;
    float *faces_dev = nullptr;        // 6·w·w·3 half-rounded floats (for rmdf_envmap_get_cube)
    cudaSurfaceObject_t surf_layered = 0, surf_cube = 0;   // the two arrays as 2D layered surfaces: the pack kernel writes the texels
};

struct rmdf_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    unsigned int *d_queue = nullptr;   // [0] work queue, [1] CTA-exit counter of the running kernel, [2] wait status
    // multi-GPU frame assembly (rmdf_frame_begin): applies to the next render / Julia launch, then clears
    unsigned int *sync_ctrl = nullptr;
    int sync_rank = 0, sync_world = 1, sync_timeout_ms = 0, sync_deal = 0;
    unsigned int sync_seq = 0;
    unsigned long long *d_counters = nullptr;
    uint64_t launches = 0;
    // work order (frame_sync.cuh): item costs of the last frame, and the permutation made from them, valid for the
    // next frame with the same work list (order_key)
    int order_mode = 1;                // rmdf_set_work_order: 0 off, 1 on
    unsigned int *d_order = nullptr, *d_rec = nullptr;
    unsigned int *d_order_sync = nullptr;   // an ordered frame's queue word [0], CTA-exit counter [1] and, from kOrderSyncHeader on, its count matrix
    size_t order_cap = 0;              // items d_order / d_rec hold
    int order_n = 0;                   // items of the last recorded frame
    bool order_sync_clean = false;     // d_order_sync is all zero (order_place_kernel leaves it so)
    int order_key[24] = {0};
    bool order_valid = false;          // d_order is a permutation for order_key
    EnvSlot env[RMDF_ENV_SLOTS];
    float4 *d_cornell = nullptr;
    bool have_cornell = false;
    void *d_out = nullptr;             // staging for the host-pointer entry points
    size_t d_out_bytes = 0;
    std::recursive_mutex mu;
    // message of the last failing call on THIS context, whatever OS thread made it (rmdf_ctx_last_error): a Haskell
    // `safe` foreign call may run the call and the error fetch on different OS threads, where the thread-local
    // rmdf_last_error() would be empty or stale
    std::mutex err_mu;
    char last_err[512] = "";
};

// Every entry point runs on ctx's device and puts the caller's current device back on return: a host
// with several GPUs in one process (or torch, which reads cudaGetDevice) must not see it change.
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    rmdf_ctx *owner = nullptr;
    int enter(rmdf_ctx *ctx)
    {
        if (!ctx) return fail(RMDF_ERR_ARG, "ctx is NULL");
        owner = ctx;
        g_err_fresh = false;
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != ctx->device) {
            CU(cudaSetDevice(ctx->device));
            changed = true;
        }
        return RMDF_OK;
    }
    ~DeviceGuard()
    {
        if (owner && g_err_fresh) {
            std::lock_guard<std::mutex> lk(owner->err_mu);
            memcpy(owner->last_err, g_err, sizeof g_err);
        }
        if (changed && prev >= 0) cudaSetDevice(prev);
    }
};
static int ensure_staging(rmdf_ctx *ctx, size_t bytes)
{
    if (ctx->d_out_bytes >= bytes) return RMDF_OK;
    if (ctx->d_out) { cudaFree(ctx->d_out); ctx->d_out = nullptr; ctx->d_out_bytes = 0; }
    CU(cudaMalloc(&ctx->d_out, bytes));
    ctx->d_out_bytes = bytes;
    return RMDF_OK;
}

// ---------------------------------------------------------------------------------
extern "C" const char *rmdf_last_error(void) { return g_err; }
extern "C" int rmdf_ctx_last_error(rmdf_ctx *ctx, char *buf, int cap)
{
    if (!ctx || !buf || cap < 1) return RMDF_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->err_mu);
    snprintf(buf, (size_t)cap, "%s", ctx->last_err);
    return RMDF_OK;
}
extern "C" int rmdf_abi_version(void) { return RMDF_ABI_VERSION; }

extern "C" int rmdf_create(rmdf_ctx **out, int device)
{
    if (!out) return fail(RMDF_ERR_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(RMDF_ERR_CUDA, "no CUDA device: %s (librmdf_b200 has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(RMDF_ERR_ARG, "device %d out of range [0,%d)", device, ndev);
    struct Restore { int prev = -1; ~Restore() { if (prev >= 0) cudaSetDevice(prev); } } restore;
    if (cudaGetDevice(&restore.prev) != cudaSuccess) { cudaGetLastError(); restore.prev = -1; }
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(RMDF_ERR_CUDA, "device %d is sm_%d%d; librmdf_b200 carries sm_100a code only", device, prop.major, prop.minor);
    rmdf_ctx *ctx = new rmdf_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t err = cudaSuccess;
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreate(&ctx->ev0);
    if (err == cudaSuccess) err = cudaEventCreate(&ctx->ev1);
    if (err == cudaSuccess) err = cudaMalloc(&ctx->d_queue, 4 * sizeof(unsigned int));
    if (err == cudaSuccess) err = cudaMemset(ctx->d_queue, 0, 4 * sizeof(unsigned int));
    if (err == cudaSuccess) err = cudaMalloc(&ctx->d_counters, sizeof(unsigned long long) * C_NSLOTS);
    if (err == cudaSuccess) err = cudaMemset(ctx->d_counters, 0, sizeof(unsigned long long) * C_NSLOTS);
    if (err == cudaSuccess) err = cudaMalloc(&ctx->d_cornell, sizeof(float4) * kCornellF4);
    if (err == cudaSuccess) err = preload_frame_sync_kernels();
    if (err == cudaSuccess) {
        // stream-ordered scratch (the convolution's partial sums) stays in the device's pool between calls instead of
        // going back to the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if (err != cudaSuccess) {
        fail(RMDF_ERR_CUDA, "rmdf_create: %s", cudaGetErrorString(err));
        rmdf_destroy(ctx);
        return RMDF_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return RMDF_OK;
}

static void free_slot(EnvSlot &s)
{
    if (s.tex_point) cudaDestroyTextureObject(s.tex_point);
    if (s.tex_cube) cudaDestroyTextureObject(s.tex_cube);
    if (s.tex_cube_nearest) cudaDestroyTextureObject(s.tex_cube_nearest);
    if (s.surf_layered) cudaDestroySurfaceObject(s.surf_layered);
    if (s.surf_cube) cudaDestroySurfaceObject(s.surf_cube);
    if (s.arr_layered) cudaFreeArray(s.arr_layered);
    if (s.arr_cube) cudaFreeArray(s.arr_cube);
    if (s.faces_dev) cudaFree(s.faces_dev);
    s = EnvSlot();
}

extern "C" void rmdf_destroy(rmdf_ctx *ctx)
{
    if (!ctx) return;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < RMDF_ENV_SLOTS; i++) free_slot(ctx->env[i]);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_cornell) cudaFree(ctx->d_cornell);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_queue) cudaFree(ctx->d_queue);
    if (ctx->d_order) cudaFree(ctx->d_order);
    if (ctx->d_rec) cudaFree(ctx->d_rec);
    if (ctx->d_order_sync) cudaFree(ctx->d_order_sync);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    if (prev >= 0) cudaSetDevice(prev);
}

extern "C" int rmdf_set_stream(rmdf_ctx *ctx, void *cuda_stream)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return RMDF_OK;
}
extern "C" int rmdf_sync(rmdf_ctx *ctx)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return RMDF_OK;
}

// ---- env maps ---------------------------------------------------------------------
extern "C" int rmdf_envmap_resize_height(int w, int h, int dstw)
{
    if (w <= 0 || h <= 0 || dstw <= 0) return 0;
    return (int)rintf((float)h / (float)w * (float)dstw);      // Haskell `round`: half to even
}

extern "C" int rmdf_envmap_resize(rmdf_ctx *ctx, const float *rgb, int w, int h, int dstw, float *out)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!rgb || !out || w < 2 || h < 2 || dstw < 1) return fail(RMDF_ERR_ARG, "rmdf_envmap_resize: bad image %dx%d -> %d", w, h, dstw);
    const int dsth = rmdf_envmap_resize_height(w, h, dstw);
    if (dsth < 1) return fail(RMDF_ERR_ARG, "rmdf_envmap_resize: target height is 0");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    float *d_src = nullptr, *d_dst = nullptr;
    const size_t nsrc = sizeof(float) * 3 * (size_t)w * h, ndst = sizeof(float) * 3 * (size_t)dstw * dsth;
    CU(cudaMalloc(&d_src, nsrc));
    cudaError_t e = cudaMalloc(&d_dst, ndst);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_src, rgb, nsrc, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = launch_resize(d_src, w, h, dstw, dsth, d_dst, ctx->stream);
    if (e == cudaSuccess) { ctx->launches++; e = cudaMemcpyAsync(out, d_dst, ndst, cudaMemcpyDeviceToHost, ctx->stream); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_src); if (d_dst) cudaFree(d_dst);
    if (e != cudaSuccess) return fail(RMDF_ERR_CUDA, "rmdf_envmap_resize: %s", cudaGetErrorString(e));
    return RMDF_OK;
}

static int convolve_check(const void *rgb, int w, int h, const float *powers, int np, const void *out)
{
    if (!rgb || !out || !powers) return fail(RMDF_ERR_ARG, "rmdf_envmap_convolve: NULL pointer");
    if (w < 2 || h < 2 || w > 1024) return fail(RMDF_ERR_ARG, "rmdf_envmap_convolve: image %dx%d unsupported (2 <= w <= 1024)", w, h);
    if (np < 1 || np > 64) return fail(RMDF_ERR_ARG, "rmdf_envmap_convolve: np = %d out of range [1,64]", np);
    return RMDF_OK;
}
extern "C" int rmdf_envmap_convolve_rows_device(rmdf_ctx *ctx, const void *rgb_dev, int w, int h,
                                                const float *powers, int np, int row_begin, int row_end, void *out_dev)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if ((rc = convolve_check(rgb_dev, w, h, powers, np, out_dev))) return rc;
    if (row_begin < 0 || row_end > h || row_begin >= row_end) return fail(RMDF_ERR_ARG, "rmdf_envmap_convolve: rows [%d,%d) outside [0,%d)", row_begin, row_end, h);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    cudaError_t e = launch_convolve((const float *)rgb_dev, w, h, powers, np, (float *)out_dev, row_begin, row_end, ctx->stream);
    if (e != cudaSuccess) return fail(RMDF_ERR_CUDA, "rmdf_envmap_convolve: %s", cudaGetErrorString(e));
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->ev_valid = true;
    ctx->launches += 1 + 2 * (uint64_t)((np + 3) / 4);          // premultiply + (convolve, finish) per batch of four powers
    return RMDF_OK;
}
extern "C" int rmdf_envmap_convolve_device(rmdf_ctx *ctx, const void *rgb_dev, int w, int h,
                                           const float *powers, int np, void *out_dev)
{
    return rmdf_envmap_convolve_rows_device(ctx, rgb_dev, w, h, powers, np, 0, h, out_dev);
}
extern "C" int rmdf_envmap_convolve(rmdf_ctx *ctx, const float *rgb, int w, int h,
                                    const float *powers, int np, float *out)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if ((rc = convolve_check(rgb, w, h, powers, np, out))) return rc;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    // source and np results in the context's grow-only staging buffer (a cudaMalloc/cudaFree pair per call cost more than the sweep)
    const size_t nsrc = sizeof(float) * 3 * (size_t)w * h;
    if ((rc = ensure_staging(ctx, nsrc * (size_t)(np + 1)))) return rc;
    float *d_src = (float *)ctx->d_out, *d_dst = d_src + 3 * (size_t)w * h;
    CU(cudaMemcpyAsync(d_src, rgb, nsrc, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = rmdf_envmap_convolve_device(ctx, d_src, w, h, powers, np, d_dst))) return rc;
    CU(cudaMemcpyAsync(out, d_dst, nsrc * np, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return RMDF_OK;
}

// JuicyPixels RGBE convention (Codec.Picture.HDR toRGBE / toFloat; not vendored in the
// reference — see DESIGN.md "third-party arithmetic").
extern "C" int rmdf_rgbe_roundtrip(float *rgb, int64_t npix)
{
    if (!rgb || npix < 0) return fail(RMDF_ERR_ARG, "rmdf_rgbe_roundtrip: bad arguments");
    for (int64_t i = 0; i < npix; i++) {
        float *p = rgb + i * 3;
        float d = p[0];
        if (p[1] > d) d = p[1];
        if (p[2] > d) d = p[2];
        if (d <= 1e-32f) { p[0] = p[1] = p[2] = 0.0f; continue; }
        int e;
        const float sig = frexpf(d, &e);
        const float coeff = sig * 255.9999f / d;
        const float f = ldexpf(1.0f, (e + 128) - (128 + 8));
        for (int c = 0; c < 3; c++) {
            const int m = (int)(unsigned char)(int)truncf(p[c] * coeff);
            p[c] = ((float)m + 0.0f) * f;
        }
    }
    return RMDF_OK;
}

// (re)fill a slot's two arrays from s.faces_dev; creates arrays + texture objects on first use
// and re-uses them when the face width is unchanged (per-frame uploads stay allocation-free)
static int make_slot_textures(rmdf_ctx *ctx, EnvSlot &s, int fw, const float *faces_dev_src)
{
    cudaStream_t st = ctx->stream;
    const bool fresh = (s.arr_layered == nullptr);
    cudaError_t e = cudaSuccess;
    const cudaChannelFormatDesc desc = cudaCreateChannelDescHalf4();
    const cudaExtent ext = make_cudaExtent((size_t)fw, (size_t)fw, 6);
    // surface-writable arrays: ONE kernel rounds the faces to half and writes the texels of both arrays (a staging buffer
    // and two cudaMemcpy3DAsync per map cost ~30 µs a map in the per-frame upload of the end-to-end path)
    if (fresh && e == cudaSuccess) e = cudaMalloc3DArray(&s.arr_layered, &desc, ext, cudaArrayLayered | cudaArraySurfaceLoadStore);
    if (fresh && e == cudaSuccess) e = cudaMalloc3DArray(&s.arr_cube, &desc, ext, cudaArrayCubemap | cudaArraySurfaceLoadStore);
    if (fresh && e == cudaSuccess) {
        cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = s.arr_layered;
        e = cudaCreateSurfaceObject(&s.surf_layered, &rd);
        rd.res.array.array = s.arr_cube;
        if (e == cudaSuccess) e = cudaCreateSurfaceObject(&s.surf_cube, &rd);
    }
    if (e == cudaSuccess) e = launch_pack_half4(faces_dev_src, fw, s.surf_layered, s.surf_cube, st);
    ctx->launches++;
    if (fresh && e == cudaSuccess) {
        cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
        rd.resType = cudaResourceTypeArray;
        cudaTextureDesc td; memset(&td, 0, sizeof td);
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.readMode = cudaReadModeElementType;
        // exact texel fetch for the parity path (FP32 weights are applied in the kernel)
        rd.res.array.array = s.arr_layered;
        td.filterMode = cudaFilterModePoint; td.normalizedCoords = 0;
        e = cudaCreateTextureObject(&s.tex_point, &rd, &td, nullptr);
        if (e == cudaSuccess) {
            // GL_TEXTURE_CUBE_MAP_SEAMLESS + LINEAR (HDREnvMap.hs:122-126) for the fast path
            rd.res.array.array = s.arr_cube;
            td.filterMode = cudaFilterModeLinear; td.normalizedCoords = 1; td.seamlessCubemap = 1;
            e = cudaCreateTextureObject(&s.tex_cube, &rd, &td, nullptr);
        }
        if (e == cudaSuccess) {
            // the same cube map with NEAREST (the GL minification filter, GLHelpers.hs:105-106): fast path of minified quads
            td.filterMode = cudaFilterModePoint; td.seamlessCubemap = 0;
            e = cudaCreateTextureObject(&s.tex_cube_nearest, &rd, &td, nullptr);
        }
    }
    if (e != cudaSuccess) {
        free_slot(s);
        return fail(RMDF_ERR_CUDA, "env-map texture setup: %s", cudaGetErrorString(e));
    }
    s.face_w = fw;
    return RMDF_OK;
}

// enqueue the upload of one slot on the stream (no synchronisation at the end)
static int bind_cube_async(rmdf_ctx *ctx, int slot, const float *faces_rgb, int face_w)
{
    if (slot < 0 || slot >= RMDF_ENV_SLOTS) return fail(RMDF_ERR_ARG, "env slot %d out of range", slot);
    if (!faces_rgb || face_w < 1 || face_w > 8192) return fail(RMDF_ERR_ARG, "rmdf_envmap_bind_cube: bad faces (w = %d)", face_w);
    EnvSlot &s = ctx->env[slot];
    const size_t bytes = sizeof(float) * 3 * 6 * (size_t)face_w * face_w;
    if (s.face_w != face_w) {                       // first bind or a size change: reallocate
        CU(cudaStreamSynchronize(ctx->stream));
        free_slot(s);
        CU(cudaMalloc(&s.faces_dev, bytes));
    }
    // stream-ordered: a render already queued on the stream keeps sampling the old texels
    cudaError_t e = cudaMemcpyAsync(s.faces_dev, faces_rgb, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { free_slot(s); return fail(RMDF_ERR_CUDA, "rmdf_envmap_bind_cube: %s", cudaGetErrorString(e)); }
    return make_slot_textures(ctx, s, face_w, s.faces_dev);
}
extern "C" int rmdf_envmap_bind_cube(rmdf_ctx *ctx, int slot, const float *faces_rgb, int face_w)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if ((rc = bind_cube_async(ctx, slot, faces_rgb, face_w))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));         // the caller's buffer is not referenced after return
    return RMDF_OK;
}
extern "C" int rmdf_envmap_bind_cubes(rmdf_ctx *ctx, int n, const int *slots, const float *const *faces_rgb, const int *face_w)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (n < 1 || n > RMDF_ENV_SLOTS || !slots || !faces_rgb || !face_w) return fail(RMDF_ERR_ARG, "rmdf_envmap_bind_cubes: bad arguments");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    for (int i = 0; i < n; i++)
        if ((rc = bind_cube_async(ctx, slots[i], faces_rgb[i], face_w[i]))) { cudaStreamSynchronize(ctx->stream); return rc; }
    CU(cudaStreamSynchronize(ctx->stream));         // one synchronisation for the batch; the buffers are free again
    return RMDF_OK;
}

extern "C" int rmdf_envmap_bind(rmdf_ctx *ctx, int slot, const float *latlong_rgb, int w, int h)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (slot < 0 || slot >= RMDF_ENV_SLOTS) return fail(RMDF_ERR_ARG, "env slot %d out of range", slot);
    if (!latlong_rgb || w < 3 || h < 2) return fail(RMDF_ERR_ARG, "rmdf_envmap_bind: bad image %dx%d", w, h);
    const int fw = w / 3;                                    // HDREnvMap.hs:128
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    CU(cudaStreamSynchronize(ctx->stream));
    free_slot(ctx->env[slot]);
    EnvSlot &s = ctx->env[slot];
    float *d_ll = nullptr;
    const size_t nll = sizeof(float) * 3 * (size_t)w * h;
    CU(cudaMalloc(&d_ll, nll));
    cudaError_t e = cudaMalloc(&s.faces_dev, sizeof(float) * 3 * 6 * (size_t)fw * fw);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ll, latlong_rgb, nll, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = launch_latlong_to_cube(d_ll, w, h, fw, s.faces_dev, ctx->stream);
    ctx->launches++;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_ll);
    if (e != cudaSuccess) { free_slot(s); return fail(RMDF_ERR_CUDA, "rmdf_envmap_bind: %s", cudaGetErrorString(e)); }
    rc = make_slot_textures(ctx, s, fw, s.faces_dev);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return RMDF_OK;
}

extern "C" int rmdf_envmap_get_cube(rmdf_ctx *ctx, int slot, float *faces_rgb, int *face_w)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (slot < 0 || slot >= RMDF_ENV_SLOTS) return fail(RMDF_ERR_ARG, "env slot %d out of range", slot);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    const EnvSlot &s = ctx->env[slot];
    if (face_w) *face_w = s.face_w;
    if (faces_rgb && s.face_w > 0) {
        CU(cudaMemcpyAsync(faces_rgb, s.faces_dev, sizeof(float) * 3 * 6 * (size_t)s.face_w * s.face_w, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return RMDF_OK;
}

extern "C" int rmdf_envmap_unbind(rmdf_ctx *ctx, int slot)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (slot < 0 || slot >= RMDF_ENV_SLOTS) return fail(RMDF_ERR_ARG, "env slot %d out of range", slot);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    CU(cudaStreamSynchronize(ctx->stream));
    free_slot(ctx->env[slot]);
    return RMDF_OK;
}

// ---- Cornell geometry --------------------------------------------------------------
// CornellBox.hs:48-129 (quads) and :26-38 (triangulation, centring, scaling)
static const float kCornellQuads[64][3] = {
    {552.8f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 559.2f}, {549.6f, 0.0f, 559.2f},
    {556.0f, 548.8f, 0.0f}, {556.0f, 548.8f, 559.2f}, {0.0f, 548.8f, 559.2f}, {0.0f, 548.8f, 0.0f},
    {549.6f, 0.0f, 559.2f}, {0.0f, 0.0f, 559.2f}, {0.0f, 548.8f, 559.2f}, {556.0f, 548.8f, 559.2f},
    {0.0f, 0.0f, 559.2f}, {0.0f, 0.0f, 0.0f}, {0.0f, 548.8f, 0.0f}, {0.0f, 548.8f, 559.2f},
    {552.8f, 0.0f, 0.0f}, {549.6f, 0.0f, 559.2f}, {556.0f, 548.8f, 559.2f}, {556.0f, 548.8f, 0.0f},
    {343.0f, -1.0f, 227.0f}, {343.0f, -1.0f, 332.0f}, {213.0f, -1.0f, 332.0f}, {213.0f, -1.0f, 227.0f},
    {130.0f, 165.0f, 65.0f}, {82.0f, 165.0f, 225.0f}, {240.0f, 165.0f, 272.0f}, {290.0f, 165.0f, 114.0f},
    {290.0f, 0.0f, 114.0f}, {290.0f, 165.0f, 114.0f}, {240.0f, 165.0f, 272.0f}, {240.0f, 0.0f, 272.0f},
    {130.0f, 0.0f, 65.0f}, {130.0f, 165.0f, 65.0f}, {290.0f, 165.0f, 114.0f}, {290.0f, 0.0f, 114.0f},
    {82.0f, 0.0f, 225.0f}, {82.0f, 165.0f, 225.0f}, {130.0f, 165.0f, 65.0f}, {130.0f, 0.0f, 65.0f},
    {240.0f, 0.0f, 272.0f}, {240.0f, 165.0f, 272.0f}, {82.0f, 165.0f, 225.0f}, {82.0f, 0.0f, 225.0f},
    {423.0f, 330.0f, 247.0f}, {265.0f, 330.0f, 296.0f}, {314.0f, 330.0f, 456.0f}, {472.0f, 330.0f, 406.0f},
    {423.0f, 0.0f, 247.0f}, {423.0f, 330.0f, 247.0f}, {472.0f, 330.0f, 406.0f}, {472.0f, 0.0f, 406.0f},
    {472.0f, 0.0f, 406.0f}, {472.0f, 330.0f, 406.0f}, {314.0f, 330.0f, 456.0f}, {314.0f, 0.0f, 456.0f},
    {314.0f, 0.0f, 456.0f}, {314.0f, 330.0f, 456.0f}, {265.0f, 330.0f, 296.0f}, {265.0f, 0.0f, 296.0f},
    {265.0f, 0.0f, 296.0f}, {265.0f, 330.0f, 296.0f}, {423.0f, 330.0f, 247.0f}, {423.0f, 0.0f, 247.0f},
};
extern "C" int rmdf_cornell_box_vertices(float *xyz)
{
    if (!xyz) return fail(RMDF_ERR_ARG, "rmdf_cornell_box_vertices: NULL");
    const float to_unit = 559.2f / 2.0f;
    const float scale = 1.0f / (sqrtf(2.0f * 2.0f + 2.0f * 2.0f + 2.0f * 2.0f) / 2.0f) * 0.99f;
    static const int corner[6] = { 0, 1, 3, 3, 1, 2 };
    int n = 0;
    for (int q = 0; q < 16; q++)
        for (int k = 0; k < 6; k++)
            for (int c = 0; c < 3; c++) {
                float v = kCornellQuads[q * 4 + corner[k]][c];
                if (v == -1.0f) v = 548.8f - 0.1f;              // light, CornellBox.hs:81-84
                xyz[n++] = (v / to_unit - 1.0f) * scale;
            }
    return RMDF_OK;
}

extern "C" int rmdf_set_cornell(rmdf_ctx *ctx, const float *xyz, int nverts)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!xyz || nverts != 96) return fail(RMDF_ERR_ARG, "rmdf_set_cornell: expected 96 vertices, got %d", nverts);
    // position-independent terms of compute_barycentric / line_seg_min_dist_sq
    // (fragment.shd:312-346), same operation order as the shader
    float4 tab[kCornellF4];
    for (int i = 0; i < 32; i++) {
        const float *t = xyz + i * 9;
        const rg_vec3 v0 = rg_v3(t[0], t[1], t[2]), v1 = rg_v3(t[3], t[4], t[5]), v2 = rg_v3(t[6], t[7], t[8]);
        const rg_vec3 e0 = rg_sub(v2, v0), e1 = rg_sub(v1, v0), ab12 = rg_sub(v2, v1);
        const float dot00 = rg_dot(e0, e0), dot01 = rg_dot(e0, e1), dot11 = rg_dot(e1, e1);
        const float inv_denom = 1.0f / (dot00 * dot11 - dot01 * dot01);
        const float len12 = rg_dot(ab12, ab12);
        float4 *r = tab + i * kCornellStride;
        r[0] = make_float4(v0.x, v0.y, v0.z, dot00);
        r[1] = make_float4(v1.x, v1.y, v1.z, dot01);
        r[2] = make_float4(v2.x, v2.y, v2.z, dot11);
        r[3] = make_float4(e0.x, e0.y, e0.z, inv_denom);
        r[4] = make_float4(e1.x, e1.y, e1.z, len12);
        r[5] = make_float4(ab12.x, ab12.y, ab12.z, 0.0f);
        // bounding sphere around the centroid, for exact culling in de_cornell_box (never enters a result)
        const rg_vec3 c = rg_v3((v0.x + v1.x + v2.x) / 3.0f, (v0.y + v1.y + v2.y) / 3.0f, (v0.z + v1.z + v2.z) / 3.0f);
        const float rad = fmaxf(rg_length(rg_sub(v0, c)), fmaxf(rg_length(rg_sub(v1, c)), rg_length(rg_sub(v2, c))));
        r[6] = make_float4(c.x, c.y, c.z, rad * 1.0001f + 1.0e-5f);
        // supporting plane (unit normal n, offset −n·v0) in double precision: |n·p + w| is a lower bound of the
        // distance to the triangle — the second exact culling test (a wall far away has a huge sphere but a far plane)
        {
            const double ax = (double)v1.x - v0.x, ay = (double)v1.y - v0.y, az = (double)v1.z - v0.z;
            const double bx = (double)v2.x - v0.x, by = (double)v2.y - v0.y, bz = (double)v2.z - v0.z;
            double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
            const double l = sqrt(nx * nx + ny * ny + nz * nz);
            if (l > 0.0) { nx /= l; ny /= l; nz /= l; } else { nx = ny = nz = 0.0; }      // degenerate: plane test never culls
            r[7] = make_float4((float)nx, (float)ny, (float)nz, (float)-(nx * v0.x + ny * v0.y + nz * v0.z));
        }
    }
    // one plane test for the two triangles of a quad (CornellBox.hs builds every quad as triangles 2k, 2k+1): the
    // pair's plane is triangle 2k's when all six vertices lie within 1e-6 of it (the culling margin is 1e-4),
    // else the null plane, whose distance 0 never culls — the per-triangle tests still follow (culling only)
    for (int k = 0; k < kCornellPairs; k++) {
        const float4 pl = tab[(2 * k) * kCornellStride + 7];
        bool coplanar = (pl.x != 0.0f || pl.y != 0.0f || pl.z != 0.0f);
        for (int j = 0; j < 6 && coplanar; j++) {
            const float *q = xyz + (2 * k) * 9 + j * 3;
            const double d = (double)pl.x * q[0] + (double)pl.y * q[1] + (double)pl.z * q[2] + (double)pl.w;
            if (!(fabs(d) <= 1.0e-6)) coplanar = false;
        }
        tab[32 * kCornellStride + k] = coplanar ? pl : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(ctx->d_cornell, tab, sizeof tab, cudaMemcpyHostToDevice));
    ctx->have_cornell = true;
    return RMDF_OK;
}

// ---- frame rendering ------------------------------------------------------------------
struct WorkList {
    int rect_mode;                    // 0 frame, 1 one tile, 2 shard
    int tile_x0, tile_y0, shard_gx, shard_skew, shard_ty0;
    int n_tiles, tile_w, tile_h;
    int64_t out_pixels;
    int out_stride;
    // rect_mode 1: the pixels that are STORED, [clip_x0,clip_x1)×[clip_y0,clip_y1).  The rendered rect
    // (tile_x0.., tile_w×tile_h) is this rectangle grown to even pixel boundaries: a 2×2 fragment quad that
    // straddles a tile edge is shaded whole (GL's helper invocations) so that its min/mag decision — and with it
    // every pixel — is the same as in the full-frame draw.
    int clip_x0, clip_y0, clip_x1, clip_y1;
    int64_t out_base;                 // index of the rendered rect's origin in the output (may be negative by the growth)
};

// First pixel whose centre lies at or beyond edge t of the reference's 8-tile split of `dim` pixels
// (ShaderRendering.hs:183-195: NDC edge −1 + t/8·2 → window coordinate t·dim/8; a pixel belongs to the tile
// that contains its centre, a centre exactly on an edge goes to the tile on its upper side): ceil(t·dim/8 − ½).
static int tile_edge(int t, int dim) { return (int)((2LL * t * dim + 7) >> 4); }

static int build_worklist(const rmdf_params *p, WorkList *wl)
{
    if (!p) return fail(RMDF_ERR_ARG, "params is NULL");
    if (p->scene < 0 || p->scene > 3) return fail(RMDF_ERR_ARG, "scene %d out of range [0,3]", p->scene);
    if (p->fb_w < 1 || p->fb_h < 1 || p->fb_w > 131072 || p->fb_h > 131072) return fail(RMDF_ERR_ARG, "frame buffer %dx%d out of range", p->fb_w, p->fb_h);
    if (p->resolve_log2 < 0 || p->resolve_log2 > 3) return fail(RMDF_ERR_ARG, "resolve_log2 %d out of range [0,3]", p->resolve_log2);
    if (p->out_format != RMDF_OUT_RGBA8 && p->out_format != RMDF_OUT_RGBA32F) return fail(RMDF_ERR_ARG, "out_format %d unknown", p->out_format);
    if (p->out_format == RMDF_OUT_RGBA32F && p->resolve_log2 != 0) return fail(RMDF_ERR_ARG, "RGBA32F output requires resolve_log2 = 0");
    if (p->mode != RMDF_MODE_PARITY && p->mode != RMDF_MODE_FAST) return fail(RMDF_ERR_ARG, "mode %d unknown", p->mode);
    if (p->lod_policy != RMDF_LOD_QUAD && p->lod_policy != RMDF_LOD_ALWAYS_LINEAR) return fail(RMDF_ERR_ARG, "lod_policy %d unknown", p->lod_policy);
    if (p->layout != RMDF_LAYOUT_FRAME && p->layout != RMDF_LAYOUT_PACKED_TILES) return fail(RMDF_ERR_ARG, "layout %d unknown", p->layout);
    if (p->flags & ~(RMDF_FLAG_NORMAL_CENTRAL | RMDF_FLAG_CAM_ORTHO | RMDF_FLAG_RAY_TRACING_TEST | RMDF_FLAG_SOFT_SHADOWS)) return fail(RMDF_ERR_ARG, "unknown bits in flags: 0x%x", p->flags);
    const int k = p->resolve_log2, q = 1 << k;
    const bool sharded = p->shard_count > 1;
    if (sharded && (p->shard_rank < 0 || p->shard_rank >= p->shard_count)) return fail(RMDF_ERR_ARG, "shard_rank %d not in [0,%d)", p->shard_rank, p->shard_count);
    if (sharded && p->tile_idx >= 0) return fail(RMDF_ERR_ARG, "tile_idx and shard_count > 1 are mutually exclusive");
    memset(wl, 0, sizeof *wl);
    const bool striped = p->row_begin != 0 || p->row_end != 0;
    if (striped) {
        if (p->row_begin < 0 || p->row_end > p->fb_h || p->row_begin >= p->row_end) return fail(RMDF_ERR_ARG, "rows [%d,%d) outside [0,%d)", p->row_begin, p->row_end, p->fb_h);
        if (p->layout != RMDF_LAYOUT_FRAME) return fail(RMDF_ERR_ARG, "a row range needs RMDF_LAYOUT_FRAME");
        if (p->tile_idx >= 0) return fail(RMDF_ERR_ARG, "tile_idx and a row range are mutually exclusive");
    }
    if (p->tile_idx < 0 && !sharded) {
        if ((p->fb_w % q) || (p->fb_h % q)) return fail(RMDF_ERR_ARG, "frame %dx%d is not a multiple of the %dx%d resolve footprint", p->fb_w, p->fb_h, q, q);
        if (p->layout != RMDF_LAYOUT_FRAME) return fail(RMDF_ERR_ARG, "packed layout needs tile_idx >= 0 or shard_count > 1");
        wl->out_stride = p->fb_w >> k;
        wl->out_pixels = (int64_t)(p->fb_w >> k) * (p->fb_h >> k);
        if (striped) {                                    // a stripe is one full-width rectangle
            if ((p->row_begin % (2 * q)) || (p->row_end % (2 * q))) return fail(RMDF_ERR_ARG, "row range must be a multiple of %d rows", 2 * q);
            wl->rect_mode = 1; wl->n_tiles = 1;
            wl->tile_x0 = 0; wl->tile_y0 = p->row_begin; wl->tile_w = p->fb_w; wl->tile_h = p->row_end - p->row_begin;
            wl->clip_x0 = 0; wl->clip_x1 = p->fb_w; wl->clip_y0 = p->row_begin; wl->clip_y1 = p->row_end;
            wl->out_base = (int64_t)(wl->tile_y0 >> k) * wl->out_stride;
            return RMDF_OK;
        }
        wl->rect_mode = 0; wl->n_tiles = 1; wl->tile_w = p->fb_w; wl->tile_h = p->fb_h;
        return RMDF_OK;
    }
    if (!sharded) {
        // one tile of the reference's 8×8 grid (idx `mod` nTiles, :186-188), any frame size
        const int t = p->tile_idx % 64, tx = t & 7, ty = t >> 3;
        wl->clip_x0 = tile_edge(tx, p->fb_w); wl->clip_x1 = tile_edge(tx + 1, p->fb_w);
        wl->clip_y0 = tile_edge(ty, p->fb_h); wl->clip_y1 = tile_edge(ty + 1, p->fb_h);
        const int cw = wl->clip_x1 - wl->clip_x0, ch = wl->clip_y1 - wl->clip_y0;
        if ((wl->clip_x0 % q) || (wl->clip_y0 % q) || (cw % q) || (ch % q))
            return fail(RMDF_ERR_ARG, "tile [%d,%d)x[%d,%d) is not a multiple of the %dx%d resolve footprint", wl->clip_x0, wl->clip_x1, wl->clip_y0, wl->clip_y1, q, q);
        wl->rect_mode = 1; wl->n_tiles = 1;
        wl->tile_x0 = wl->clip_x0 & ~1; wl->tile_y0 = wl->clip_y0 & ~1;
        const int ex1 = (wl->clip_x1 + 1) & ~1, ey1 = (wl->clip_y1 + 1) & ~1;
        wl->tile_w = (ex1 < p->fb_w ? ex1 : p->fb_w) - wl->tile_x0;
        wl->tile_h = (ey1 < p->fb_h ? ey1 : p->fb_h) - wl->tile_y0;
        if (cw <= 0 || ch <= 0) { wl->tile_w = wl->tile_h = 0; }
        if (p->layout == RMDF_LAYOUT_PACKED_TILES) {
            wl->out_stride = cw >> k;
            wl->out_pixels = (int64_t)(cw >> k) * (ch >> k);
            wl->out_base = -((int64_t)((wl->clip_y0 - wl->tile_y0) >> k) * wl->out_stride + ((wl->clip_x0 - wl->tile_x0) >> k));
        } else {
            wl->out_stride = p->fb_w >> k;
            wl->out_pixels = (int64_t)(p->fb_w >> k) * (p->fb_h >> k);
            wl->out_base = (int64_t)(wl->tile_y0 >> k) * wl->out_stride + (wl->tile_x0 >> k);
        }
        return RMDF_OK;
    }
    // Shard grids (multi-GPU; not a reference concept): tile edges must fall on even pixel boundaries so that no
    // 2×2 fragment quad straddles two tiles.
    int gx = p->shard_grid_x ? p->shard_grid_x : 8;
    int gy = p->shard_grid_y ? p->shard_grid_y : 8;
    wl->shard_skew = p->shard_grid_x ? p->shard_skew : 1;
    if (gx < 1 || gy < 1 || gx > 1024 || gy > 1024) return fail(RMDF_ERR_ARG, "shard grid %dx%d out of range", gx, gy);
    if (gx % p->shard_count) return fail(RMDF_ERR_ARG, "shard_grid_x %d is not a multiple of shard_count %d", gx, p->shard_count);
    if (wl->shard_skew < 0) return fail(RMDF_ERR_ARG, "shard_skew must be >= 0");
    if ((p->fb_w % (2 * gx)) || (p->fb_h % (2 * gy)))
        return fail(RMDF_ERR_UNSUPPORTED, "sharded rendering needs frame dimensions divisible by %dx%d (got %dx%d)", 2 * gx, 2 * gy, p->fb_w, p->fb_h);
    wl->tile_w = p->fb_w / gx; wl->tile_h = p->fb_h / gy;
    if ((wl->tile_w % q) || (wl->tile_h % q)) return fail(RMDF_ERR_ARG, "tile %dx%d is not a multiple of the %dx%d resolve footprint", wl->tile_w, wl->tile_h, q, q);
    {
        int ty0 = 0, ty1 = gy;
        if (striped) {
            if ((p->row_begin % wl->tile_h) || (p->row_end % wl->tile_h)) return fail(RMDF_ERR_ARG, "row range must be a multiple of the shard tile height %d", wl->tile_h);
            ty0 = p->row_begin / wl->tile_h; ty1 = p->row_end / wl->tile_h;
        }
        wl->rect_mode = 2; wl->shard_gx = gx; wl->shard_ty0 = ty0;
        wl->n_tiles = (ty1 - ty0) * (gx / p->shard_count);
    }
    if (p->layout == RMDF_LAYOUT_PACKED_TILES) {
        wl->out_stride = wl->tile_w >> k;
        wl->out_pixels = (int64_t)wl->n_tiles * (wl->tile_w >> k) * (wl->tile_h >> k);
    } else {
        wl->out_stride = p->fb_w >> k;
        wl->out_pixels = (int64_t)(p->fb_w >> k) * (p->fb_h >> k);
    }
    return RMDF_OK;
}

// ---- work order (frame_sync.cuh) --------------------------------------------------------------------------------
// order_attach: point the frame's FrameSync at the cost record (always, when the work list qualifies) and at the
// permutation of the previous frame (when that frame had the same work list); order_finish: after the frame's
// kernel, turn its record into the permutation for the next frame.
static int order_min_items()
{
    static const int v = [] { const char *e = getenv("RMDF_ORDER_MIN_ITEMS"); return e ? atoi(e) : 4096; }();
    return v;
}
static bool order_attach(rmdf_ctx *ctx, const int key[24], int n_items, FrameSync *fs)
{
    fs->order = nullptr; fs->rec = nullptr; fs->counts = nullptr;
    if (!ctx->order_mode || fs->queue_sys || n_items < order_min_items()) return false;
    if (ctx->order_cap < (size_t)n_items) {
        // cudaFree waits for the device: no frame that still uses the old buffers is in flight afterwards
        if (ctx->d_order) { cudaFree(ctx->d_order); ctx->d_order = nullptr; }
        if (ctx->d_rec) { cudaFree(ctx->d_rec); ctx->d_rec = nullptr; }
        if (ctx->d_order_sync) { cudaFree(ctx->d_order_sync); ctx->d_order_sync = nullptr; }
        ctx->order_cap = 0; ctx->order_valid = false;
        const size_t chunks = ((size_t)n_items + kOrderChunk - 1) / kOrderChunk;
        if (cudaMalloc(&ctx->d_order, (size_t)n_items * sizeof(unsigned int)) != cudaSuccess ||
            cudaMalloc(&ctx->d_rec, (size_t)n_items * sizeof(unsigned int)) != cudaSuccess ||
            cudaMalloc(&ctx->d_order_sync, (kOrderSyncHeader + chunks * kOrderBuckets) * sizeof(unsigned int)) != cudaSuccess) {
            cudaGetLastError();                              // no memory for hints: render in scan order
            if (ctx->d_order) { cudaFree(ctx->d_order); ctx->d_order = nullptr; }
            if (ctx->d_rec) { cudaFree(ctx->d_rec); ctx->d_rec = nullptr; }
            return false;
        }
        ctx->order_cap = (size_t)n_items;
        ctx->order_sync_clean = false;
    }
    if (memcmp(key, ctx->order_key, sizeof ctx->order_key) != 0) {
        memcpy(ctx->order_key, key, sizeof ctx->order_key);
        ctx->order_valid = false;
    }
    // an ordered frame has its queue word, its CTA-exit counter and its count matrix in ONE buffer: one memset clears them
    if (fs->queue == ctx->d_queue) fs->queue = ctx->d_order_sync;
    fs->done = ctx->d_order_sync + 1;
    fs->order = ctx->order_valid ? ctx->d_order : nullptr;
    fs->rec = ctx->d_rec;
    fs->counts = ctx->d_order_sync + kOrderSyncHeader;
    return true;
}
// Clears what the frame's kernel counts in: the queue word and the CTA-exit counter and, for an ordered frame, its count
// matrix — which the previous ordered frame's order_place_kernel has normally already done (order_sync_clean).
static cudaError_t clear_frame_sync(rmdf_ctx *ctx, const FrameSync &fs)
{
    if (fs.counts == nullptr) return cudaMemsetAsync(ctx->d_queue, 0, 2 * sizeof(unsigned int), ctx->stream);
    if (ctx->order_sync_clean) return cudaSuccess;
    const size_t chunks = (ctx->order_cap + kOrderChunk - 1) / kOrderChunk;
    return cudaMemsetAsync(ctx->d_order_sync, 0, (kOrderSyncHeader + chunks * kOrderBuckets) * sizeof(unsigned int), ctx->stream);
}
static int order_finish(rmdf_ctx *ctx, const FrameSync &fs, int n_items)
{
    ctx->order_n = n_items;
    ctx->order_sync_clean = false;
    if (getenv("RMDF_ORDER_DEBUG")) return RMDF_OK;         // experiments: the host sorts (rmdf_debug_order_*)
    cudaError_t e = launch_order_place(fs.rec, (unsigned)n_items, ctx->d_order_sync, ctx->d_order, ctx->stream);
    if (e != cudaSuccess) { ctx->order_valid = false; return fail(RMDF_ERR_CUDA, "work-order kernel launch: %s", cudaGetErrorString(e)); }
    ctx->order_valid = true;
    ctx->order_sync_clean = true;                           // its last CTA zeroes the control buffer
    ctx->launches++;
    return RMDF_OK;
}
// experiments only (profiles/exp_order.py): the cost record of the last frame / a permutation for the next one
extern "C" int rmdf_debug_order_get(rmdf_ctx *ctx, unsigned int *rec_host, int cap)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!ctx->d_rec || ctx->order_n <= 0 || cap < ctx->order_n) return -1;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(rec_host, ctx->d_rec, (size_t)ctx->order_n * 4, cudaMemcpyDeviceToHost));
    return ctx->order_n;
}
extern "C" int rmdf_debug_order_get_order(rmdf_ctx *ctx, unsigned int *order_host, int cap)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!ctx->d_order || !ctx->order_valid || ctx->order_n <= 0 || cap < ctx->order_n) return -1;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(order_host, ctx->d_order, (size_t)ctx->order_n * 4, cudaMemcpyDeviceToHost));
    return ctx->order_n;
}
extern "C" int rmdf_debug_order_set(rmdf_ctx *ctx, const unsigned int *order_host, int n)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!ctx->d_order || n != ctx->order_n) return -1;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(ctx->d_order, order_host, (size_t)n * 4, cudaMemcpyHostToDevice));
    ctx->order_valid = true;
    return 0;
}
extern "C" int rmdf_set_work_order(rmdf_ctx *ctx, int enabled)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    ctx->order_mode = enabled ? 1 : 0;
    ctx->order_valid = false;
    return RMDF_OK;
}

extern "C" int64_t rmdf_render_output_pixels(const rmdf_params *p, int32_t *n_tiles)
{
    WorkList wl;
    if (build_worklist(p, &wl)) return -1;
    if (n_tiles) *n_tiles = wl.n_tiles;
    return wl.out_pixels;
}

// orbit camera + lookat (fragment.shd:829-838, 884-902), hoisted: it is constant per frame
static void make_camera(int scene, float in_time, float cam[12])
{
    rg_vec3 cam_pos;
    if (scene == RMDF_SCENE_CORNELL_BOX) {
        cam_pos.x = rg_sin(in_time / 2.0f) * 0.4f;
        cam_pos.z = -2.0f;
        cam_pos.y = rg_cos(in_time / 2.0f) * 0.4f;
    } else {
        cam_pos.x = rg_sin(in_time / 3.0f);
        cam_pos.z = rg_cos(in_time / 3.0f);
        cam_pos.y = rg_cos(in_time / 4.0f);
        cam_pos = rg_scale(rg_normalize(cam_pos), 2.414213562373095f);
    }
    const rg_vec3 zaxis = rg_normalize(rg_sub(cam_pos, rg_v3(0.0f, 0.0f, 0.0f)));
    const rg_vec3 xaxis = rg_normalize(rg_cross(rg_v3(0.0f, 1.0f, 0.0f), zaxis));
    const rg_vec3 yaxis = rg_cross(zaxis, xaxis);
    cam[0] = xaxis.x; cam[1] = xaxis.y; cam[2] = xaxis.z;
    cam[3] = yaxis.x; cam[4] = yaxis.y; cam[5] = yaxis.z;
    cam[6] = zaxis.x; cam[7] = zaxis.y; cam[8] = zaxis.z;
    cam[9] = cam_pos.x; cam[10] = cam_pos.y; cam[11] = cam_pos.z;
}

static int pick_item_h(int64_t blocks_of_32_rows, int resident_warps, int k)
{
    // shrink the work item until there are ≥ 16 items per resident warp: small frames fill the
    // machine, and a rank's share of a sharded frame ends without a long tail
    int item_h = 32;
    const int min_h = (1 << k) > 4 ? (1 << k) : 4;
    while (item_h > min_h && blocks_of_32_rows * (32 / item_h) < (int64_t)resident_warps * 16) item_h >>= 1;
    return item_h;
}

extern "C" int rmdf_render_device(rmdf_ctx *ctx, const rmdf_params *p, void *fb_dev)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    WorkList wl;
    if ((rc = build_worklist(p, &wl))) return rc;
    if (!fb_dev) return fail(RMDF_ERR_ARG, "frame buffer pointer is NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if ((p->scene == RMDF_SCENE_CORNELL_BOX || (p->flags & RMDF_FLAG_RAY_TRACING_TEST)) && !ctx->have_cornell)
        return fail(RMDF_ERR_STATE, "Cornell scene requested before rmdf_set_cornell");

    MarchParams mp;
    memset(&mp, 0, sizeof mp);
    mp.scene = p->scene; mp.fb_w = p->fb_w; mp.fb_h = p->fb_h;
    mp.fw = (float)p->fb_w; mp.fh = (float)p->fb_h;
    mp.aspect = mp.fw / mp.fh;
    mp.fov_xs = rg_tan(rg_radians(45.0f * 1.5f) / 2.0f);               // fragment.shd:866-867,910
    make_camera(p->scene, p->time, mp.cam);
    {                                                                   // fragment.shd:116-119
        float pow_offs = rg_mod(p->time / 2.0f, 9.0f);
        if (pow_offs > 4.5f) pow_offs = 9.0f - pow_offs;
        mp.power = pow_offs + 2.0f;
    }
    mp.bsphere_r = (p->scene == RMDF_SCENE_MB_POWER8) ? 1.15f : (p->scene == RMDF_SCENE_MB_GENERAL) ? 1.5f : 1.0f;
    mp.lod_policy = p->lod_policy; mp.resolve_log2 = p->resolve_log2;
    mp.out_format = p->out_format; mp.layout = p->layout; mp.flags = p->flags;
    mp.rect_mode = wl.rect_mode; mp.tile_x0 = wl.tile_x0; mp.tile_y0 = wl.tile_y0;
    mp.clip_x0 = wl.clip_x0; mp.clip_y0 = wl.clip_y0; mp.clip_x1 = wl.clip_x1; mp.clip_y1 = wl.clip_y1;
    mp.rect_out_base = wl.out_base;
    mp.clipped = (wl.rect_mode == 1) && (wl.clip_x0 != wl.tile_x0 || wl.clip_y0 != wl.tile_y0 ||
                                         wl.clip_x1 != wl.tile_x0 + wl.tile_w || wl.clip_y1 != wl.tile_y0 + wl.tile_h);
    mp.shard_gx = wl.shard_gx; mp.shard_count = p->shard_count; mp.shard_rank = p->shard_rank; mp.shard_skew = wl.shard_skew;
    mp.shard_ty0 = wl.shard_ty0;
    mp.tile_w = wl.tile_w; mp.tile_h = wl.tile_h;
    mp.out_stride = wl.out_stride;
    for (int s = 0; s < RMDF_ENV_SLOTS; s++) {
        mp.env_point[s] = ctx->env[s].tex_point;
        mp.env_cube[s] = ctx->env[s].tex_cube;
        mp.env_cube_nearest[s] = ctx->env[s].tex_cube_nearest;
        mp.env_face[s] = ctx->env[s].face_w;
    }
    mp.cornell = ctx->d_cornell;
    mp.out = fb_dev;
    mp.sync.queue = ctx->d_queue; mp.sync.done = ctx->d_queue + 1;
    mp.sync.signal = nullptr; mp.sync.signal_value = 0; mp.sync.queue_sys = 0;
    mp.sync.item_stride = 1; mp.sync.item_offset = 0; mp.sync.item_group = 1; mp.sync.items_total = 0;
    int world = 1;
    unsigned int *go_word = nullptr;
    if (ctx->sync_ctrl) {                                   // armed by rmdf_frame_begin: one frame over several GPUs
        if (wl.rect_mode != 0 || p->layout != RMDF_LAYOUT_FRAME)
            return fail(RMDF_ERR_ARG, "a multi-GPU frame (rmdf_frame_begin) needs the full frame in RMDF_LAYOUT_FRAME (no tile, shard or row range)");
        if (ctx->sync_deal == RMDF_DEAL_QUEUE) { mp.sync.queue = ctx->sync_ctrl + kSyncQueueWord; mp.sync.queue_sys = 1; }
        else { mp.sync.item_stride = (unsigned)ctx->sync_world; mp.sync.item_offset = (unsigned)ctx->sync_rank; }
        mp.sync.signal = ctx->sync_ctrl + kSyncFlagWord0 + ctx->sync_rank; mp.sync.signal_value = ctx->sync_seq;
        go_word = ctx->sync_ctrl + kSyncGoWord;
        world = ctx->sync_world;
        ctx->sync_ctrl = nullptr;
    }
    mp.counters = ctx->d_counters;

    const int ctas_per_sm = (p->mode == RMDF_MODE_FAST) ? march_fast_ctas_per_sm(p->scene, p->flags) : march_parity_ctas_per_sm(p->scene, p->flags);
    const int resident_warps = ctas_per_sm * ctx->sm_count * kWarpsPerCta * world;     // all GPUs that draw from the queue
    // Work-item size.  Measured on BASELINE config 4 (profiles/r2_item_size.md): 32×32-sample items 17.34 ms, 32×16
    // 16.86, 32×8 16.70, 32×4 16.74, 16×4 17.11 — small items interleave the per-item resolve/store bursts of the
    // warps and keep the end of the frame short, very small ones pay the claim/decode/store overhead more often.  So
    // items start at 32×8 and shrink (rows to 4, then columns to 16 and 8: one 8×4 tile row) while there are fewer
    // than 24 items per resident warp: mid-size frames (config 3, 3840×2160) land on 16×4, config 1 (512²) on 8×4.
    int bx;
    {
        const int q = 1 << p->resolve_log2;
        const int min_h = q > 4 ? q : 4, min_w = q > 8 ? q : 8;
        mp.item_h = 8 > min_h ? 8 : min_h; mp.item_w = kBlockDim;
        auto items = [&]() { return (int64_t)((wl.tile_w + mp.item_w - 1) / mp.item_w) * ((wl.tile_h + mp.item_h - 1) / mp.item_h) * wl.n_tiles; };
        // 32×8 -> 32×4 already below 100 items per warp: the two cost the same per sample (16.70 / 16.74 ms), but the kernel's
        // idle end is one worst-case item long — 8 tiles walked by one warp, ~0.3 ms (per-warp end times, profiles/dbg_tail.py:
        // mean idle 0.23 ms with 32×8, 0.11 ms with 32×4) — which weighs 2×, 4× as much on a half, a quarter of the frame
        if (mp.item_h > min_h && mp.item_h == 8 && items() < (int64_t)resident_warps * 100) mp.item_h >>= 1;
        while (items() < (int64_t)resident_warps * 24) {
            if (mp.item_h > min_h) mp.item_h >>= 1;
            else if (mp.item_w > min_w) mp.item_w >>= 1;
            else break;
        }
        bx = (wl.tile_w + mp.item_w - 1) / mp.item_w;
    }
    if (const char *e = getenv("RMDF_ITEM_H")) {          // tuning overrides (experiments only)
        const int v = atoi(e), min_h = (1 << p->resolve_log2) > 4 ? (1 << p->resolve_log2) : 4;
        if (v >= min_h && v <= 32 && (v & (v - 1)) == 0) mp.item_h = v;
    }
    if (const char *e = getenv("RMDF_ITEM_W")) {
        const int v = atoi(e), min_w = (1 << p->resolve_log2) > 8 ? (1 << p->resolve_log2) : 8;
        if (v >= min_w && v <= 32 && (v & (v - 1)) == 0) { mp.item_w = v; bx = (wl.tile_w + mp.item_w - 1) / mp.item_w; }
    }
    mp.blocks_x = bx;
    mp.blocks_y = (wl.tile_h + mp.item_h - 1) / mp.item_h;
    // reciprocal for decode_item: q = (n·inv) >> 40 with inv = floor(2^40/d) + 1 equals floor(n/d) while n·d < 2^40
    // (error n·(inv − 2^40/d)/2^40 <= n/2^40 < 1/d); n < blocks_x·blocks_y
    mp.blocks_x_inv = (mp.blocks_x > 0 && (unsigned long long)mp.blocks_x * mp.blocks_y * (unsigned long long)mp.blocks_x < (1ull << 40))
                    ? ((1ull << 40) / (unsigned long long)mp.blocks_x + 1ull) : 0ull;
    const int64_t n_items = (int64_t)mp.blocks_x * mp.blocks_y * wl.n_tiles;
    if (n_items > 0x7fffffff) return fail(RMDF_ERR_UNSUPPORTED, "frame too large: %lld work items", (long long)n_items);
    mp.n_items = (int)n_items;
    mp.sync.items_total = (unsigned)n_items;
    if (mp.sync.item_stride > 1) {
        // interleaved deal: groups of x-adjacent items that cover >= 128 B of an output row go round-robin to the ranks;
        // the kernel walks this rank's groups as local items 0 .. n-1 (frame_sync.cuh frame_item)
        const int out_bytes_per_item_row = (mp.item_w >> p->resolve_log2) * ((p->out_format == RMDF_OUT_RGBA32F) ? 16 : 4);
        int g = 1;
        while (g * out_bytes_per_item_row < 128 && g * 2 <= mp.blocks_x) g *= 2;
        mp.sync.item_group = (unsigned)g;
        const int64_t n_groups = (n_items + g - 1) / g;
        // one group per rank out of every window of `world` consecutive groups (the slot rotates: frame_sync.cuh frame_item);
        // in the last, partial window some ranks' slot lies past the frame (the kernel skips it)
        const int64_t mine = (n_groups + mp.sync.item_stride - 1) / mp.sync.item_stride;
        mp.sync.item_rotate = getenv("RMDF_NO_DEAL_ROTATE") ? 0u : 1u;
        mp.n_items = (int)(mine * g);
    }
    if (mp.n_items == 0 && mp.sync.signal == nullptr) return RMDF_OK;
    const int okey[24] = { 1, p->scene, p->fb_w, p->fb_h, wl.rect_mode, wl.tile_x0, wl.tile_y0, wl.tile_w, wl.tile_h, wl.n_tiles,
                           p->shard_count, p->shard_rank, wl.shard_skew, wl.shard_ty0, mp.item_w, mp.item_h, mp.n_items,
                           (int)mp.sync.item_stride, (int)mp.sync.item_offset, (int)mp.sync.item_group, p->resolve_log2, p->flags, (int)mp.sync.item_rotate, 0 };
    // work order: the fractals have the heavy-tailed item costs it pays for (measured: config 4 −0.35 %, one GPU's eighth of it
    // −2.5 %, general power −5 %, config 1 −1…4 %); the Cornell box gains nothing and the CSG scene loses the 10 µs it costs
    // (the other scenes' kernels are compiled without it: march_kernel kOrdered)
    // (a frame drawn as the reference's 64 tiles cycles through 64 work lists: nothing to learn from the call before)
    const bool placed = (p->scene == RMDF_SCENE_MB_POWER8 || p->scene == RMDF_SCENE_MB_GENERAL);
    const bool ordered = placed && wl.rect_mode != 1 && order_attach(ctx, okey, mp.n_items, &mp.sync);
    if (!ordered) { mp.sync.order = nullptr; mp.sync.rec = nullptr; mp.sync.counts = nullptr; }
    int grid = ctas_per_sm * ctx->sm_count;
    // A frame of very few items is bound by the latency of its slowest one (config 1: one 8×4 tile marches for 80 of the
    // frame's 90 µs), and that item runs faster with fewer warps on its scheduler: at least four items per warp
    // (config 1: 8 → 3 CTAs per SM, 0.091 → 0.078 ms fast, 0.217 → 0.204 ms parity)
    {
        const int cap = (mp.n_items / 4 + kWarpsPerCta - 1) / kWarpsPerCta;
        if (grid > cap && !getenv("RMDF_NO_GRID_CAP")) grid = cap > ctx->sm_count ? cap : (grid < ctx->sm_count ? grid : ctx->sm_count);
    }
    if (const char *e = getenv("RMDF_GRID_CTAS_PER_SM")) {   // tuning override (experiments only)
        const int v = atoi(e);
        if (v >= 1 && v < ctas_per_sm) grid = v * ctx->sm_count;
    }
    const int need = (mp.n_items + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid > need) grid = need > 0 ? need : 1;             // a rank without items still publishes its arrival flag
    mp.sync.first_wave = (!placed || mp.sync.queue_sys || getenv("RMDF_NO_FIRST_WAVE")) ? 0u : (unsigned)(grid * kWarpsPerCta);

    if (go_word && ctx->sync_rank != 0) {                  // rank != 0 of a multi-GPU frame: wait for rank 0's "go" (see rmdf_frame_begin)
        cudaError_t ew = launch_wait_flags(go_word, 1, mp.sync.signal_value, (unsigned)ctx->sync_timeout_ms,
                                           ctx->d_queue + 2, ctx->stream);
        if (ew != cudaSuccess) return fail(RMDF_ERR_CUDA, "frame wait launch: %s", cudaGetErrorString(ew));
        ctx->launches++;
    }
    CU(clear_frame_sync(ctx, mp.sync));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    cudaError_t e = (p->mode == RMDF_MODE_FAST) ? launch_march_fast(mp, grid, ctx->stream)
                                                : launch_march_parity(mp, grid, ctx->stream);
    if (e != cudaSuccess) return fail(RMDF_ERR_CUDA, "march kernel launch: %s", cudaGetErrorString(e));
    ctx->launches++;
    if (ordered && (rc = order_finish(ctx, mp.sync, mp.n_items))) return rc;   // inside the timed span: it is part of the frame's cost
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->ev_valid = true;
    return RMDF_OK;
}

// A host buffer the device can write directly: page-locked memory (cudaHostAlloc / cudaHostRegister — e.g. a
// pinned Storable vector or torch pin_memory) is mapped into the device's address space under UVA.  The kernels then
// store resolved pixels straight into it over PCIe while they run, and the separate device→host copy of the frame
// (33 MB for BASELINE config 4: 18.2 → 17.7 ms end to end) disappears.  Pageable memory → NULL.
static void *device_view_of_host(const void *host)
{
    if (getenv("RMDF_NO_ZERO_COPY")) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (at.type == cudaMemoryTypeHost) ? at.devicePointer : nullptr;
}

extern "C" int rmdf_render(rmdf_ctx *ctx, const rmdf_params *p, void *fb_host)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    WorkList wl;
    if ((rc = build_worklist(p, &wl))) return rc;
    if (!fb_host) return fail(RMDF_ERR_ARG, "frame buffer pointer is NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    const size_t px_bytes = (p->out_format == RMDF_OUT_RGBA32F) ? 16 : 4;
    const size_t bytes = (size_t)wl.out_pixels * px_bytes;
    if (void *direct = device_view_of_host(fb_host)) {     // page-locked frame buffer: the kernel's stores are the copy
        if ((rc = rmdf_render_device(ctx, p, direct))) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
        return RMDF_OK;
    }
    if ((rc = ensure_staging(ctx, bytes))) return rc;
    const bool partial = wl.rect_mode != 0 && p->layout == RMDF_LAYOUT_FRAME;
    if (partial)   // drawShaderTile leaves the rest of the FBO untouched: start from the caller's pixels
        CU(cudaMemcpyAsync(ctx->d_out, fb_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = rmdf_render_device(ctx, p, ctx->d_out))) return rc;
    CU(cudaMemcpyAsync(fb_host, ctx->d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return RMDF_OK;
}

// ---- Julia ---------------------------------------------------------------------------------
static int julia_params(rmdf_ctx *ctx, int w, int h, int smooth, double tick, int y0, int y1, int k,
                        int band_h, int shard_rank, int shard_count, JuliaParams *jp)
{
    if (w < 1 || h < 1 || w > 131072 || h > 131072) return fail(RMDF_ERR_ARG, "julia frame %dx%d out of range", w, h);
    if (k < 0 || k > 3) return fail(RMDF_ERR_ARG, "resolve_log2 %d out of range [0,3]", k);
    const int q = 1 << k;
    if (y0 < 0 || y1 > h || y0 >= y1) return fail(RMDF_ERR_ARG, "julia rows [%d,%d) outside [0,%d)", y0, y1, h);
    if ((w % q) || (y0 % q) || (y1 % q)) return fail(RMDF_ERR_ARG, "julia width/rows must be multiples of %d", q);
    if (shard_count < 1) shard_count = 1;
    if (shard_rank < 0 || shard_rank >= shard_count) return fail(RMDF_ERR_ARG, "shard_rank %d not in [0,%d)", shard_rank, shard_count);
    if (shard_count == 1 && band_h <= 0) band_h = y1 - y0;
    if (band_h <= 0 || (band_h % q) || ((y1 - y0) % band_h)) return fail(RMDF_ERR_ARG, "band height %d must divide the %d rows and be a multiple of %d", band_h, y1 - y0, q);
    const int n_bands = (y1 - y0) / band_h;
    if (n_bands % shard_count) return fail(RMDF_ERR_ARG, "%d bands do not split evenly over %d ranks", n_bands, shard_count);
    memset(jp, 0, sizeof *jp);
    // Fractal2D.hs:65-76 (host, libm like GHC's Float primops)
    const float pi = 3.14159265358979323846f;
    const float f_tick = (float)tick;
    const float q1 = f_tick / 17.0f, q2 = f_tick / 61.0f, q3 = f_tick / 71.0f;
    const float st1 = q1 - truncf(q1), st2 = q2 - truncf(q2), st3 = q3 - truncf(q3);
    const float two_pi = st1 * 2.0f * pi;
    jp->julia_r = sinf(two_pi) * (0.7f > st2 ? 0.7f : st2);
    jp->julia_i = cosf(two_pi) * (0.7f > st3 ? 0.7f : st3);
    jp->w = w; jp->h = h; jp->y0 = y0; jp->y1 = y1; jp->smooth = smooth ? 1 : 0; jp->resolve_log2 = k;
    jp->band_h = band_h; jp->shard_rank = shard_rank; jp->shard_count = shard_count;
    jp->fw = (float)w; jp->fh = (float)h;
    jp->ratio = jp->fw / jp->fh;
    jp->rcp_fw = 1.0f / jp->fw; jp->rcp_fh = 1.0f / jp->fh;      // host code is built without fast math: correctly rounded
    {
        uint32_t bw, bh;
        memcpy(&bw, &jp->fw, 4); memcpy(&bh, &jp->fh, 4);
        jp->known_div = ((bw & 0x7fffffu) != 0x7fffffu) && ((bh & 0x7fffffu) != 0x7fffffu);
    }
    jp->xshift = 1.45f * jp->ratio;
    const int my_bands = n_bands / shard_count;
    const int ctas = julia_ctas_per_sm();
    const int resident_warps = ctas * ctx->sm_count * kWarpsPerCta * (ctx->sync_ctrl ? ctx->sync_world : 1);
    const int bx = (w + kBlockDim - 1) / kBlockDim;
    jp->item_h = pick_item_h((int64_t)bx * ((band_h + 31) / 32) * my_bands, resident_warps, k);
    if (const char *e = getenv("RMDF_JULIA_ITEM_H")) {    // tuning override (experiments only)
        const int v = atoi(e), min_h = (1 << k) > 4 ? (1 << k) : 4;
        if (v >= min_h && v <= 32 && (v & (v - 1)) == 0 && band_h % v == 0) jp->item_h = v;
    }
    jp->blocks_x = bx;
    jp->blocks_y = (band_h + jp->item_h - 1) / jp->item_h;
    const int64_t n = (int64_t)jp->blocks_x * jp->blocks_y * my_bands;
    if (n > 0x7fffffff) return fail(RMDF_ERR_UNSUPPORTED, "julia frame too large");
    jp->n_items = (int)n;
    jp->out_stride = w >> k;
    jp->sync.queue = ctx->d_queue; jp->sync.done = ctx->d_queue + 1;
    jp->sync.signal = nullptr; jp->sync.signal_value = 0; jp->sync.queue_sys = 0;
    jp->sync.item_stride = 1; jp->sync.item_offset = 0; jp->sync.item_group = 1; jp->sync.items_total = 0;
    jp->counters = ctx->d_counters;
    return RMDF_OK;
}

static int julia_launch(rmdf_ctx *ctx, JuliaParams &jp, void *fb_dev)
{
    jp.out = (uint32_t *)fb_dev;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    jp.sync.items_total = (unsigned)jp.n_items;
    unsigned int *go_word = nullptr;
    if (ctx->sync_ctrl) {                                   // armed by rmdf_frame_begin: one frame over several GPUs
        if (jp.shard_count != 1 || !jp.frame_layout || jp.y0 != 0 || jp.y1 != jp.h) {
            ctx->sync_ctrl = nullptr;
            return fail(RMDF_ERR_ARG, "a multi-GPU Julia frame (rmdf_frame_begin) needs the whole frame, one band list, frame layout");
        }
        if (ctx->sync_deal == RMDF_DEAL_QUEUE) { jp.sync.queue = ctx->sync_ctrl + kSyncQueueWord; jp.sync.queue_sys = 1; }
        else {
            jp.sync.item_stride = (unsigned)ctx->sync_world; jp.sync.item_offset = (unsigned)ctx->sync_rank;
            int g = 1;                                      // groups of x-adjacent 32-sample items covering >= 128 B of an output row
            while (g * ((kBlockDim >> jp.resolve_log2) * 4) < 128 && g * 2 <= jp.blocks_x) g *= 2;
            jp.sync.item_group = (unsigned)g;
            const int n_groups = (jp.n_items + g - 1) / g;
            const int mine = (n_groups + ctx->sync_world - 1) / ctx->sync_world;
            jp.sync.item_rotate = 0u;                       // the Julia frame's fixed-slot deal is even to 0.4 % already; rotated it is 1 % slower (locality)
            jp.n_items = mine * g;
        }
        jp.sync.signal = ctx->sync_ctrl + kSyncFlagWord0 + ctx->sync_rank; jp.sync.signal_value = ctx->sync_seq;
        go_word = ctx->sync_ctrl + kSyncGoWord;
        ctx->sync_ctrl = nullptr;
    }
    if (jp.n_items == 0 && jp.sync.signal == nullptr) return RMDF_OK;
    jp.sync.order = nullptr; jp.sync.rec = nullptr; jp.sync.counts = nullptr;   // no work order: 2 M cheap items, an order costs 3 % (profiles/r2_work_order.md)
    int grid = julia_ctas_per_sm() * ctx->sm_count;
    const int need = (jp.n_items + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid > need) grid = need > 0 ? need : 1;
    jp.sync.first_wave = 0u;
    if (go_word && ctx->sync_rank != 0) {                  // rank != 0 of a multi-GPU frame: wait for rank 0's "go"
        cudaError_t ew = launch_wait_flags(go_word, 1, jp.sync.signal_value, (unsigned)ctx->sync_timeout_ms,
                                           ctx->d_queue + 2, ctx->stream);
        if (ew != cudaSuccess) return fail(RMDF_ERR_CUDA, "frame wait launch: %s", cudaGetErrorString(ew));
        ctx->launches++;
    }
    CU(clear_frame_sync(ctx, jp.sync));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    cudaError_t e = launch_julia(jp, grid, ctx->stream);
    if (e != cudaSuccess) return fail(RMDF_ERR_CUDA, "julia kernel launch: %s", cudaGetErrorString(e));
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->ev_valid = true;
    return RMDF_OK;
}

extern "C" int rmdf_julia_device(rmdf_ctx *ctx, int w, int h, int smooth, double tick,
                                 int y0, int y1, int resolve_log2, void *fb_dev)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!fb_dev) return fail(RMDF_ERR_ARG, "frame buffer pointer is NULL");
    JuliaParams jp;
    if ((rc = julia_params(ctx, w, h, smooth, tick, y0, y1, resolve_log2, 0, 0, 1, &jp))) return rc;
    jp.frame_layout = 1;                                    // one band = the row range itself: both layouts coincide
    return julia_launch(ctx, jp, fb_dev);
}

extern "C" int rmdf_julia_bands_device(rmdf_ctx *ctx, int w, int h, int smooth, double tick, int resolve_log2,
                                       int band_h, int shard_rank, int shard_count, int layout, void *fb_dev)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!fb_dev) return fail(RMDF_ERR_ARG, "frame buffer pointer is NULL");
    if (layout != RMDF_LAYOUT_FRAME && layout != RMDF_LAYOUT_PACKED_TILES) return fail(RMDF_ERR_ARG, "layout %d unknown", layout);
    JuliaParams jp;
    if ((rc = julia_params(ctx, w, h, smooth, tick, 0, h, resolve_log2, band_h, shard_rank, shard_count, &jp))) return rc;
    jp.frame_layout = (layout == RMDF_LAYOUT_FRAME) ? 1 : 0;
    return julia_launch(ctx, jp, fb_dev);
}

extern "C" int rmdf_julia(rmdf_ctx *ctx, int w, int h, int smooth, double tick,
                          int y0, int y1, int resolve_log2, uint32_t *fb_host)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!fb_host) return fail(RMDF_ERR_ARG, "frame buffer pointer is NULL");
    if (resolve_log2 < 0 || resolve_log2 > 3 || y1 <= y0 || w < 1) return fail(RMDF_ERR_ARG, "rmdf_julia: bad arguments");
    const size_t bytes = sizeof(uint32_t) * (size_t)(w >> resolve_log2) * (size_t)((y1 - y0) >> resolve_log2);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    // (no direct stores into page-locked memory here, unlike rmdf_render: the Julia kernel emits 25 GB/s of 16-byte
    //  row fragments — measured on config 5: 10.4 ms direct against 7.7 ms through the staging buffer + one copy)
    if ((rc = ensure_staging(ctx, bytes))) return rc;
    if ((rc = rmdf_julia_device(ctx, w, h, smooth, tick, y0, y1, resolve_log2, ctx->d_out))) return rc;
    CU(cudaMemcpyAsync(fb_host, ctx->d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return RMDF_OK;
}

// ---- frame-buffer helpers ----------------------------------------------------------------
extern "C" int rmdf_mip_resolve_device(rmdf_ctx *ctx, const void *fb_dev, int w, int h, int levels, void *out_dev)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!fb_dev || !out_dev || w < 1 || h < 1 || levels < 0 || levels > 16 || (w >> levels) < 1 || (h >> levels) < 1)
        return fail(RMDF_ERR_ARG, "rmdf_mip_resolve_device: bad arguments");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    cudaError_t e = launch_mip_resolve((const uint32_t *)fb_dev, w, h, levels, (uint32_t *)out_dev, ctx->stream);
    if (e != cudaSuccess) return fail(RMDF_ERR_CUDA, "mip resolve: %s", cudaGetErrorString(e));
    ctx->launches += (uint64_t)(levels > 0 ? levels : 0);
    return RMDF_OK;
}

// ---- multi-GPU shared frame buffer ------------------------------------------------------------
extern "C" int rmdf_shared_alloc(rmdf_ctx *ctx, int64_t bytes, void **dev_ptr, unsigned char handle[64])
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!dev_ptr || !handle || bytes <= 0) return fail(RMDF_ERR_ARG, "rmdf_shared_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    CU(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);        // a control block inside starts at sequence number 0
    if (e != cudaSuccess) { cudaFree(p); return fail(RMDF_ERR_CUDA, "rmdf_shared_alloc: %s", cudaGetErrorString(e)); }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(RMDF_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle, &h, 64);
    *dev_ptr = p;
    return RMDF_OK;
}
extern "C" int rmdf_shared_free(rmdf_ctx *ctx, void *dev_ptr)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(dev_ptr));
    return RMDF_OK;
}
extern "C" int rmdf_shared_open(rmdf_ctx *ctx, const unsigned char handle[64], void **dev_ptr)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!dev_ptr || !handle) return fail(RMDF_ERR_ARG, "rmdf_shared_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RMDF_OK;
}
extern "C" int rmdf_shared_close(rmdf_ctx *ctx, void *dev_ptr)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return RMDF_OK;
}
extern "C" int rmdf_enable_peer(rmdf_ctx *ctx, int peer_device)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) return fail(RMDF_ERR_UNSUPPORTED, "device %d cannot access device %d", ctx->device, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    if (e != cudaSuccess) return fail(RMDF_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
    return RMDF_OK;
}

extern "C" int rmdf_host_register(rmdf_ctx *ctx, void *host_ptr, int64_t bytes, void **dev_ptr)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!host_ptr || !dev_ptr || bytes <= 0) return fail(RMDF_ERR_ARG, "rmdf_host_register: bad arguments");
    cudaError_t e = cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(RMDF_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); }
    e = cudaHostGetDevicePointer(dev_ptr, host_ptr, 0);
    if (e != cudaSuccess) { cudaHostUnregister(host_ptr); cudaGetLastError(); return fail(RMDF_ERR_CUDA, "cudaHostGetDevicePointer: %s", cudaGetErrorString(e)); }
    return RMDF_OK;
}
extern "C" int rmdf_host_unregister(rmdf_ctx *ctx, void *host_ptr)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!host_ptr) return fail(RMDF_ERR_ARG, "rmdf_host_unregister: NULL");
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaHostUnregister(host_ptr));
    return RMDF_OK;
}

// ---- one frame over several GPUs: shared work queue + arrival flags in peer memory (csrc/frame_sync.cuh) --------
extern "C" int rmdf_frame_begin(rmdf_ctx *ctx, void *ctrl_dev, int rank, int world, uint32_t seq, int deal, int timeout_ms)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!ctrl_dev || world < 1 || world > 32 || rank < 0 || rank >= world) return fail(RMDF_ERR_ARG, "rmdf_frame_begin: bad arguments (rank %d of %d)", rank, world);
    if (deal != RMDF_DEAL_INTERLEAVED && deal != RMDF_DEAL_QUEUE) return fail(RMDF_ERR_ARG, "rmdf_frame_begin: deal %d unknown", deal);
    if (timeout_ms <= 0) timeout_ms = 10000;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    unsigned int *ctrl = (unsigned int *)ctrl_dev;
    if (rank == 0) {                                       // queue = 0, then go = seq
        cudaError_t e = launch_frame_begin(ctrl, seq, ctx->stream);
        if (e != cudaSuccess) return fail(RMDF_ERR_CUDA, "rmdf_frame_begin: %s", cudaGetErrorString(e));
        ctx->launches++;
    }
    // The other ranks wait for "go" on their stream — but that wait kernel is enqueued by the render call itself, after
    // it has made sure its own kernel is loaded: CUDA loads kernels lazily and a first load can need the device to
    // drain, which a spinning wait kernel would hold up until its time-out.
    ctx->sync_ctrl = ctrl; ctx->sync_rank = rank; ctx->sync_world = world; ctx->sync_seq = seq; ctx->sync_timeout_ms = timeout_ms; ctx->sync_deal = deal;
    return RMDF_OK;
}
extern "C" int rmdf_frame_end(rmdf_ctx *ctx, void *ctrl_dev, int world, uint32_t seq, int timeout_ms)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!ctrl_dev || world < 1 || world > 32) return fail(RMDF_ERR_ARG, "rmdf_frame_end: bad arguments");
    if (timeout_ms <= 0) timeout_ms = 10000;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    cudaError_t e = launch_wait_flags((unsigned int *)ctrl_dev + kSyncFlagWord0, world, seq, (unsigned)timeout_ms, ctx->d_queue + 2, ctx->stream);
    if (e != cudaSuccess) return fail(RMDF_ERR_CUDA, "rmdf_frame_end: %s", cudaGetErrorString(e));
    ctx->launches++;
    return RMDF_OK;
}
extern "C" int rmdf_frame_sync_status(rmdf_ctx *ctx, int *timed_out_word)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!timed_out_word) return fail(RMDF_ERR_ARG, "rmdf_frame_sync_status: NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    unsigned int v = 0;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(&v, ctx->d_queue + 2, sizeof v, cudaMemcpyDeviceToHost));
    CU(cudaMemset(ctx->d_queue + 2, 0, sizeof v));
    *timed_out_word = (int)v;
    return RMDF_OK;
}

// ---- measurement ----------------------------------------------------------------------------
extern "C" int rmdf_measure_fp32_peak(rmdf_ctx *ctx, float *tflops, float *sm_mhz)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!tflops || !sm_mhz) return fail(RMDF_ERR_ARG, "rmdf_measure_fp32_peak: NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    cudaError_t e = measure_fp32_peak(ctx->sm_count, ctx->stream, tflops, sm_mhz);
    if (e != cudaSuccess) return fail(RMDF_ERR_CUDA, "rmdf_measure_fp32_peak: %s", cudaGetErrorString(e));
    ctx->launches += 6;
    return RMDF_OK;
}

extern "C" int rmdf_measure_mufu_peak(rmdf_ctx *ctx, float *gops)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!gops) return fail(RMDF_ERR_ARG, "rmdf_measure_mufu_peak: NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    cudaError_t e = measure_mufu_peak(ctx->sm_count, ctx->stream, gops);
    if (e != cudaSuccess) return fail(RMDF_ERR_CUDA, "rmdf_measure_mufu_peak: %s", cudaGetErrorString(e));
    ctx->launches += 5;
    return RMDF_OK;
}

// ---- counters -----------------------------------------------------------------------------
extern "C" int rmdf_counters_get(rmdf_ctx *ctx, rmdf_counters *out)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    if (!out) return fail(RMDF_ERR_ARG, "out is NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    unsigned long long h[C_NSLOTS];
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof *out);
    out->samples = h[C_SAMPLES]; out->in_sphere = h[C_IN_SPHERE]; out->hits = h[C_HITS];
    out->march_de = h[C_MARCH_DE]; out->shade_de = h[C_SHADE_DE]; out->inner_iters = h[C_ITERS];
    out->launches = ctx->launches;
    if (ctx->ev_valid) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) out->last_kernel_ms = ms;
    }
    return RMDF_OK;
}
extern "C" int rmdf_counters_reset(rmdf_ctx *ctx)
{
    DeviceGuard guard__; int rc = guard__.enter(ctx); if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    CU(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long) * C_NSLOTS, ctx->stream));
    return RMDF_OK;
}
