/* oracle_envmap.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's Haskell environment-map preprocessing:
 *   HDREnvMap.hs   :31-52 load, :55-73 test map, :76-87 cubeMapPixelToDir,
 *                  :91-113 pixelAtBilinear, :118-163 lat-long → cube map,
 *                  :169-195 resizeHDRImage, :217-254 cosineConvolveHDREnvMap
 *   CoordTransf.hs :27-91 coordinate conventions
 *   CornellBox.hs  :21-129 geometry table
 *   ConcurrentSegments.hs :14-28 makeNSegments
 * Haskell `Float` arithmetic = IEEE binary32; `cos sin acos atan2 **` on Float are
 * GHC primops that call libm's cosf/sinf/acosf/atan2f/powf, so this file calls the
 * same libm functions (glibc here).  Compile with -ffp-contract=off.
 *
 * PARITY STATUS: "parity unpinned" (see oracle_march.c).  Two third-party pieces
 * are not in /root/reference and are restated from their published behaviour:
 *   - JuicyPixels (rmdf.cabal: >= 3.1.7; Stackage lts-4.0) Codec.Picture.HDR: the
 *     RGBE <-> float convention.  decode: (mantissa + RGBE_DECODE_BIAS)·2^(e-136);
 *     encode: d = max(r,g,b); d <= 1e-32 → (0,0,0,0); else
 *     trunc(v · significand(d)·255.9999/d), e = exponent(d)+128 (Haskell
 *     significand ∈ [0.5,1)).
 *   - linear (>= 1.10.1.1) Linear.Metric.normalize: returns v unchanged when the
 *     squared length l has |l| <= 1e-6 or |1-l| <= 1e-6, else v / sqrt l.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/rmdf_glsl.h"   /* only rg_round_to_half (GL RGB16F upload) */
#include "oracle.h"

/* JuicyPixels' toFloat adds 0.0 (not Radiance's 0.5) to the mantissa — from memory
 * of Codec/Picture/HDR.hs; unverifiable offline. */
#define RGBE_DECODE_BIAS 0.0f

static const float PI_F = 3.14159265358979323846f;   /* Haskell `pi :: Float` */

/* ---- Radiance .hdr (JuicyPixels decodeHDR / encodeHDR) ---------------------- */
static void rgbe_to_float(const uint8_t p[4], float out[3])
{
    float f = ldexpf(1.0f, (int)p[3] - (128 + 8));
    out[0] = ((float)p[0] + RGBE_DECODE_BIAS) * f;
    out[1] = ((float)p[1] + RGBE_DECODE_BIAS) * f;
    out[2] = ((float)p[2] + RGBE_DECODE_BIAS) * f;
}
static void float_to_rgbe(const float in[3], uint8_t p[4])
{
    float d = in[0];
    int e;
    float sig, coeff;
    if (in[1] > d) d = in[1];
    if (in[2] > d) d = in[2];
    if (d <= 1e-32f) { p[0] = p[1] = p[2] = p[3] = 0; return; }
    sig = frexpf(d, &e);                       /* significand ∈ [0.5,1), exponent */
    coeff = sig * 255.9999f / d;
    p[0] = (uint8_t)(int)truncf(in[0] * coeff);
    p[1] = (uint8_t)(int)truncf(in[1] * coeff);
    p[2] = (uint8_t)(int)truncf(in[2] * coeff);
    p[3] = (uint8_t)(e + 128);
}
void oracle_rgbe_roundtrip(const float *rgb, int64_t npix, float *out)
{
    int64_t i;
    for (i = 0; i < npix; i++) {
        uint8_t p[4];
        float_to_rgbe(rgb + i * 3, p);
        rgbe_to_float(p, out + i * 3);
    }
}
/* returns 0 on success; rgb row 0 = top scanline (-Y h +X w) */
int oracle_hdr_decode(const uint8_t *file, int64_t nbytes, int *w_out, int *h_out, float *rgb, int64_t cap_floats)
{
    int64_t pos = 0;
    int w = 0, h = 0, y, x, c;
    uint8_t *line;
    /* header: lines until an empty one, then the resolution line */
    for (;;) {
        int64_t start = pos;
        while (pos < nbytes && file[pos] != '\n') pos++;
        if (pos >= nbytes) return 1;
        pos++;
        if (pos - start == 1) break;
    }
    {
        char res[64];
        int64_t start = pos, n;
        while (pos < nbytes && file[pos] != '\n') pos++;
        if (pos >= nbytes) return 1;
        n = pos - start; if (n > 63) n = 63;
        memcpy(res, file + start, (size_t)n); res[n] = 0;
        pos++;
        if (sscanf(res, "-Y %d +X %d", &h, &w) != 2) return 2;
    }
    *w_out = w; *h_out = h;
    if (!rgb) return 0;
    if ((int64_t)w * h * 3 > cap_floats) return 3;
    line = (uint8_t *)malloc((size_t)w * 4);
    for (y = 0; y < h; y++) {
        if (pos + 4 <= nbytes && w >= 8 && w < 32768 && file[pos] == 2 && file[pos + 1] == 2 &&
            ((file[pos + 2] << 8) | file[pos + 3]) == w) {
            pos += 4;                                    /* new-style RLE, per channel */
            for (c = 0; c < 4; c++) {
                x = 0;
                while (x < w) {
                    int cnt;
                    if (pos >= nbytes) { free(line); return 4; }
                    cnt = file[pos++];
                    if (cnt > 128) {
                        cnt -= 128;
                        if (pos >= nbytes || x + cnt > w) { free(line); return 4; }
                        while (cnt--) line[(x++) * 4 + c] = file[pos];
                        pos++;
                    } else {
                        if (pos + cnt > nbytes || x + cnt > w) { free(line); return 4; }
                        while (cnt--) line[(x++) * 4 + c] = file[pos++];
                    }
                }
            }
        } else {
            if (pos + (int64_t)w * 4 > nbytes) { free(line); return 4; }
            memcpy(line, file + pos, (size_t)w * 4);
            pos += (int64_t)w * 4;
        }
        for (x = 0; x < w; x++) rgbe_to_float(line + x * 4, rgb + ((size_t)y * w + x) * 3);
    }
    free(line);
    return 0;
}
int64_t oracle_hdr_encode(const float *rgb, int w, int h, uint8_t *out, int64_t cap)
{
    char hdr[128];
    int n = snprintf(hdr, sizeof hdr, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", h, w);
    int64_t need = n + (int64_t)w * h * 4, i;
    if (!out || cap < need) return need;
    memcpy(out, hdr, (size_t)n);
    for (i = 0; i < (int64_t)w * h; i++) float_to_rgbe(rgb + i * 3, out + n + i * 4);
    return need;
}

/* ---- HDREnvMap.hs:91-113 ------------------------------------------------------ */
static void pixel_at_bilinear(const float *img, int w, int h, float u, float v, float out[3])
{
    float upx = u * ((float)w - 1.0f);
    float upy = v * ((float)h - 1.0f);
    int x = (int)floorf(upx);
    int y = (int)floorf(upy);
    int xp1 = (x + 1) % (w - 1);          /* sic: mod (w-1), HDREnvMap.hs:100 */
    int yp1 = (y + 1 < h - 1) ? y + 1 : h - 1;
    float u_ratio = upx - (float)x, v_ratio = upy - (float)y;
    float u_opp = 1.0f - u_ratio, v_opp = 1.0f - v_ratio;
    const float *t00 = img + ((size_t)y * w + x) * 3, *t10 = img + ((size_t)y * w + xp1) * 3;
    const float *t01 = img + ((size_t)yp1 * w + x) * 3, *t11 = img + ((size_t)yp1 * w + xp1) * 3;
    int c;
    for (c = 0; c < 3; c++)
        out[c] = (t00[c] * u_opp + t10[c] * u_ratio) * v_opp + (t01[c] * u_opp + t11[c] * u_ratio) * v_ratio;
}

/* ---- HDREnvMap.hs:169-195 ------------------------------------------------------ */
int oracle_resize_height(int srcw, int srch, int dstw)
{
    float r = (float)srch / (float)srcw * (float)dstw;
    return (int)rintf(r);                 /* Haskell round = half-to-even = rintf */
}
void oracle_resize(const float *src, int srcw, int srch, int dstw, float *dst)
{
    int dsth = oracle_resize_height(srcw, srch, dstw);
    float scale = (float)srcw / (float)dstw;
    int taps = (int)ceilf(scale);
    float ntaps = (float)(taps * taps);
    float step = scale / (float)taps;
    int dstx, dsty, tx, ty;
    for (dsty = 0; dsty < dsth; dsty++)
        for (dstx = 0; dstx < dstw; dstx++) {
            float srcx1 = (float)dstx * scale, srcy1 = (float)dsty * scale;
            float acc[3] = { 0.0f, 0.0f, 0.0f };
            for (ty = 0; ty < taps; ty++)
                for (tx = 0; tx < taps; tx++) {
                    float srcx = srcx1 + (float)tx * step, srcy = srcy1 + (float)ty * step;
                    float u = srcx / ((float)srcw - 1.0f), v = srcy / ((float)srch - 1.0f);
                    float p[3];
                    pixel_at_bilinear(src, srcw, srch, u, v, p);
                    acc[0] = acc[0] + p[0]; acc[1] = acc[1] + p[1]; acc[2] = acc[2] + p[2];
                }
            dst[((size_t)dsty * dstw + dstx) * 3 + 0] = acc[0] / ntaps;
            dst[((size_t)dsty * dstw + dstx) * 3 + 1] = acc[1] / ntaps;
            dst[((size_t)dsty * dstw + dstx) * 3 + 2] = acc[2] / ntaps;
        }
}

/* ---- HDREnvMap.hs:217-254  O(n^4) cosine-power convolution -------------------- */
/* destination rows [y0,y1) only (the CPU-baseline sample of bench.py); dst is a full-size image */
void oracle_convolve_rows(const float *src, int srcw, int srch, float power, float *dst, int y0, int y1, int threads)
{
    int dsty;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (dsty = y0; dsty < y1; dsty++) {
        float *lookup = (float *)malloc(sizeof(float) * (size_t)srcw);
        int dstx, x, y;
        for (dstx = 0; dstx < srcw; dstx++) {
            float theta_lobe = (float)dsty / (float)(srch - 1) * PI_F;
            float theta_lobe_cos = cosf(theta_lobe), theta_lobe_sin = sinf(theta_lobe);
            float phi_lobe = (float)dstx / (float)(srcw - 1) * 2.0f * PI_F;
            float ar = 0.0f, ag = 0.0f, ab = 0.0f, n = 0.0f;
            for (x = 0; x < srcw; x++)
                lookup[x] = cosf(fabsf(phi_lobe - (float)x / (float)(srcw - 1) * 2.0f * PI_F));
            for (y = 0; y < srch; y++) {
                float theta_px = (float)y / (float)(srch - 1) * PI_F;
                float theta_px_cos = cosf(theta_px), theta_px_sin = sinf(theta_px);
                const float *row = src + (size_t)y * srcw * 3;
                for (x = 0; x < srcw; x++) {
                    float cos_angle = theta_lobe_cos * theta_px_cos + theta_lobe_sin * theta_px_sin * lookup[x];
                    if (cos_angle > 0.0f) {
                        float fac = theta_px_sin * powf(cos_angle, power);
                        ar = ar + (row[x * 3] * fac);
                        ag = ag + (row[x * 3 + 1] * fac);
                        ab = ab + (row[x * 3 + 2] * fac);
                        n = n + 1.0f;
                    }
                }
            }
            dst[((size_t)dsty * srcw + dstx) * 3 + 0] = ar / n;
            dst[((size_t)dsty * srcw + dstx) * 3 + 1] = ag / n;
            dst[((size_t)dsty * srcw + dstx) * 3 + 2] = ab / n;
        }
        free(lookup);
    }
}
void oracle_convolve(const float *src, int srcw, int srch, float power, float *dst, int threads)
{
    oracle_convolve_rows(src, srcw, srch, power, dst, 0, srch, threads);
}
/* The reference's own decomposition of the sweep (ShaderRendering.hs:142 `mapConcurrently` over the
 * powers, each cosineConvolveHDREnvMap sequential): one thread per power; rows [y0,y1) of every image. */
void oracle_convolve_sweep_rows(const float *src, int srcw, int srch, const float *powers, int np, float *dst, int y0, int y1)
{
    int k;
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(np)
#endif
    for (k = 0; k < np; k++)
        oracle_convolve_rows(src, srcw, srch, powers[k], dst + (size_t)k * srcw * srch * 3, y0, y1, 1);
}

/* ---- linear's normalize (see header) ------------------------------------------ */
static void linear_normalize(float v[3])
{
    float l = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
    float s;
    if (fabsf(l) <= 1e-6f || fabsf(1.0f - l) <= 1e-6f) return;
    s = sqrtf(l);
    v[0] = v[0] / s; v[1] = v[1] / s; v[2] = v[2] / s;
}

/* ---- HDREnvMap.hs:76-87 -------------------------------------------------------- */
void oracle_cube_pixel_to_dir(int face, int w, int x, int y, float d[3])
{
    float vw = ((float)x + 0.5f) / (float)w * 2.0f - 1.0f;
    float vh = ((float)y + 0.5f) / (float)w * 2.0f - 1.0f;
    switch (face) {
    case 0:  d[0] = 1.0f;  d[1] = -vh;   d[2] = -vw;   break;   /* +X */
    case 1:  d[0] = -1.0f; d[1] = -vh;   d[2] = vw;    break;   /* -X */
    case 2:  d[0] = vw;    d[1] = 1.0f;  d[2] = vh;    break;   /* +Y */
    case 3:  d[0] = vw;    d[1] = -1.0f; d[2] = -vh;   break;   /* -Y */
    case 4:  d[0] = vw;    d[1] = -vh;   d[2] = 1.0f;  break;   /* +Z */
    default: d[0] = -vw;   d[1] = -vh;   d[2] = -1.0f; break;   /* -Z */
    }
    linear_normalize(d);
}

/* ---- CoordTransf.hs:35-45 (cartesianToSpherical), :61-71 (sphericalToEnvironmentUV) */
static void cartesian_to_spherical(const float d[3], float *theta, float *phi)
{
    float z = d[2] < 1.0f ? d[2] : 1.0f;           /* clamp v mi ma = max mi (min ma v) */
    float p;
    z = z > -1.0f ? z : -1.0f;
    *theta = acosf(z);
    p = atan2f(d[1], d[0]);
    if (p < 0.0f) p = p + 2.0f * PI_F;
    if (p == 2.0f * PI_F) p = 0.0f;
    *phi = p;
}
static void spherical_to_environment_uv(float theta, float phi, float *u, float *v)
{
    float p1 = phi + PI_F / 2.0f;
    float p2 = (p1 > 2.0f * PI_F) ? p1 - 2.0f * PI_F : p1;
    float p3 = 2.0f * PI_F - p2;
    *u = p3 / (PI_F * 2.0f);
    *v = theta / PI_F;
}

/* ---- HDREnvMap.hs:118-163 ------------------------------------------------------- */
void oracle_latlong_to_cube(const float *latlong, int w, int h, float *faces, int round_half)
{
    int fw = w / 3, face, x, y, c;
    for (face = 0; face < 6; face++)
        for (y = 0; y < fw; y++)
            for (x = 0; x < fw; x++) {
                float dir[3], local[3], theta, phi, u, v, col[3];
                float *o = faces + (((size_t)face * fw + y) * fw + x) * 3;
                oracle_cube_pixel_to_dir(face, fw, x, y, dir);
                /* worldToLocal, CoordTransf.hs:47-51: dots with x=(1,0,0) y=(0,0,-1) n=(0,1,0) */
                local[0] = (dir[0] * 1.0f + dir[1] * 0.0f) + dir[2] * 0.0f;
                local[1] = (dir[0] * 0.0f + dir[1] * 0.0f) + dir[2] * -1.0f;
                local[2] = (dir[0] * 0.0f + dir[1] * 1.0f) + dir[2] * 0.0f;
                cartesian_to_spherical(local, &theta, &phi);
                spherical_to_environment_uv(theta, phi, &u, &v);
                pixel_at_bilinear(latlong, w, h, u, v, col);
                for (c = 0; c < 3; c++) o[c] = round_half ? rg_round_to_half(col[c]) : col[c];
            }
}

/* ---- HDREnvMap.hs:55-73 + CoordTransf.hs:27-33,53-59,81-91 ---------------------- */
void oracle_test_latlong_envmap(float *rgb)
{
    const int w = 512, h = 256;
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) {
            /* environmentPxToSpherical x y w */
            float u = (float)x / (float)(w - 1), v = (float)y / (float)(w / 2 - 1);
            float theta = v * PI_F;
            float p2 = u * PI_F * 2.0f + PI_F / 2.0f;
            float p1 = (p2 >= PI_F * 2.0f) ? p2 - PI_F * 2.0f : p2;
            float phi = 2.0f * PI_F - p1;
            float lx = sinf(theta) * cosf(phi), ly = sinf(theta) * sinf(phi), lz = cosf(theta);
            /* localToWorld: columns x=(1,0,0) y=(0,0,-1) n=(0,1,0) */
            float wx = (lx * 1.0f + ly * 0.0f) + lz * 0.0f;
            float wy = (lx * 0.0f + ly * 0.0f) + lz * 1.0f;
            float wz = (lx * 0.0f + ly * -1.0f) + lz * 0.0f;
            float ax = fabsf(wx), ay = fabsf(wy), az = fabsf(wz);
            float *o = rgb + ((size_t)y * w + x) * 3;
            if (ax >= ay && ax >= az) {
                if (wx > 0.0f) { o[0] = 1; o[1] = 0; o[2] = 0; } else { o[0] = 0; o[1] = 1; o[2] = 0; }
            } else if (ay >= ax && ay >= az) {
                if (wy > 0.0f) { o[0] = 0; o[1] = 0; o[2] = 1; } else { o[0] = 1; o[1] = 0; o[2] = 1; }
            } else {
                if (wz < 0.0f) { o[0] = 1; o[1] = 1; o[2] = 0; } else { o[0] = 0; o[1] = 1; o[2] = 1; }
            }
        }
}

/* ---- CornellBox.hs:26-38 (scaling) and :48-129 (table) -------------------------- */
static const float cornell_quads[64][3] = {
    /* floor */        {552.8f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 559.2f}, {549.6f, 0.0f, 559.2f},
    /* ceiling */      {556.0f, 548.8f, 0.0f}, {556.0f, 548.8f, 559.2f}, {0.0f, 548.8f, 559.2f}, {0.0f, 548.8f, 0.0f},
    /* back wall */    {549.6f, 0.0f, 559.2f}, {0.0f, 0.0f, 559.2f}, {0.0f, 548.8f, 559.2f}, {556.0f, 548.8f, 559.2f},
    /* right wall */   {0.0f, 0.0f, 559.2f}, {0.0f, 0.0f, 0.0f}, {0.0f, 548.8f, 0.0f}, {0.0f, 548.8f, 559.2f},
    /* left wall */    {552.8f, 0.0f, 0.0f}, {549.6f, 0.0f, 559.2f}, {556.0f, 548.8f, 559.2f}, {556.0f, 548.8f, 0.0f},
    /* light, y = 548.8 - 0.1 evaluated in Float below */
                       {343.0f, -1.0f, 227.0f}, {343.0f, -1.0f, 332.0f}, {213.0f, -1.0f, 332.0f}, {213.0f, -1.0f, 227.0f},
    /* short block */  {130.0f, 165.0f, 65.0f}, {82.0f, 165.0f, 225.0f}, {240.0f, 165.0f, 272.0f}, {290.0f, 165.0f, 114.0f},
                       {290.0f, 0.0f, 114.0f}, {290.0f, 165.0f, 114.0f}, {240.0f, 165.0f, 272.0f}, {240.0f, 0.0f, 272.0f},
                       {130.0f, 0.0f, 65.0f}, {130.0f, 165.0f, 65.0f}, {290.0f, 165.0f, 114.0f}, {290.0f, 0.0f, 114.0f},
                       {82.0f, 0.0f, 225.0f}, {82.0f, 165.0f, 225.0f}, {130.0f, 165.0f, 65.0f}, {130.0f, 0.0f, 65.0f},
                       {240.0f, 0.0f, 272.0f}, {240.0f, 165.0f, 272.0f}, {82.0f, 165.0f, 225.0f}, {82.0f, 0.0f, 225.0f},
    /* tall block */   {423.0f, 330.0f, 247.0f}, {265.0f, 330.0f, 296.0f}, {314.0f, 330.0f, 456.0f}, {472.0f, 330.0f, 406.0f},
                       {423.0f, 0.0f, 247.0f}, {423.0f, 330.0f, 247.0f}, {472.0f, 330.0f, 406.0f}, {472.0f, 0.0f, 406.0f},
                       {472.0f, 0.0f, 406.0f}, {472.0f, 330.0f, 406.0f}, {314.0f, 330.0f, 456.0f}, {314.0f, 0.0f, 456.0f},
                       {314.0f, 0.0f, 456.0f}, {314.0f, 330.0f, 456.0f}, {265.0f, 330.0f, 296.0f}, {265.0f, 0.0f, 296.0f},
                       {265.0f, 0.0f, 296.0f}, {265.0f, 330.0f, 296.0f}, {423.0f, 330.0f, 247.0f}, {423.0f, 0.0f, 247.0f},
};
void oracle_cornell_vertices(float *xyz)
{
    const float to_unit = 559.2f / 2.0f;
    const float scale = 1.0f / (sqrtf(2.0f * 2.0f + 2.0f * 2.0f + 2.0f * 2.0f) / 2.0f) * 0.99f;
    static const int corner[6] = { 0, 1, 3, 3, 1, 2 };     /* CornellBox.hs:32-37 */
    int q, k, c, n = 0;
    for (q = 0; q < 16; q++)
        for (k = 0; k < 6; k++) {
            const float *v = cornell_quads[q * 4 + corner[k]];
            for (c = 0; c < 3; c++) {
                float val = v[c];
                if (val == -1.0f) val = 548.8f - 0.1f;      /* light height, :81-84 */
                xyz[n++] = (val / to_unit - 1.0f) * scale;
            }
        }
}

/* ---- ConcurrentSegments.hs:14-23 ------------------------------------------------ */
int oracle_make_segments(int nseg, int low, int high, int *out, int cap_pairs)
{
    int range = high - low, nsegc, segl, i, n = 0;
    if (low >= high) return 0;
    nsegc = nseg < range ? nseg : range;
    if (nsegc <= 0) return 0;
    if (nsegc == 1) { if (cap_pairs > 0) { out[0] = low; out[1] = high; } return 1; }
    segl = range / nsegc;
    for (i = 0; i <= nsegc - 2; i++, n++)
        if (n < cap_pairs) { out[n * 2] = low + i * segl; out[n * 2 + 1] = low + (i + 1) * segl; }
    if (n < cap_pairs) { out[n * 2] = low + (nsegc - 1) * segl; out[n * 2 + 1] = high; }
    return n + 1;
}
