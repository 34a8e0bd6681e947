/* ref_driver.cpp — harness around the reference's OWN fragment shader, compiled by g++.
 *
 * TEST INFRASTRUCTURE (oracle/_ref recipe; see oracle/Makefile target `ref`).  This file
 * contains none of the reference's algorithm: it #includes the shader text generated from
 * /root/reference/fragment.shd (oracle/ref/gen_shader.py — `out` parameters become C++
 * references, nothing else changes) on top of oracle/ref/glsl_compat.hpp, once per
 * translation unit, with the `#define`s ShaderRendering.hs:119-123 prepends for the variant:
 *
 *   REF_VARIANT 0  CORNELL_BOX_SCENE           (FSDECornellBoxShader)
 *               1  (none)                      (FSDETestShader)
 *               2  MANDELBULB_SCENE + POWER8   (FSMBPower8Shader)
 *               3  MANDELBULB_SCENE            (FSMBGeneralShader)
 *               4  0 + RAY_TRACING_TEST        (fragment.shd:914, commented out in the reference)
 *               5  1 + CAM_ORTHO               (fragment.shd:905, idem)
 *               6  2 + CAM_ORTHO
 *
 * What the harness supplies is what OpenGL supplies around a fragment shader: uniforms,
 * gl_FragCoord at pixel centres (row 0 = bottom), the 2×2-quad min/mag decision for
 * texture() (conventions C4/C5 — the sampler itself is the oracle's, GL fixed function is
 * not in the reference's source), float → RGBA8 conversion of frag_color (C6).
 */
#include <stdint.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../../include/rmdf_glsl.h"
#include "../oracle.h"

extern "C" {
/* exported by liboracle.so (oracle_march.c): the GL cube sampler, conventions C4/C5 */
void oracle_texture_cube(const oracle_frame *f, int slot, const float d[3], int nearest, float rgb[3]);
int  oracle_quad_is_minified(int face_w, const float d[12]);
}

#ifndef REF_VARIANT
#error "REF_VARIANT must be defined"
#endif
#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)
#define REF_NAME(x) REF_CAT(REF_CAT(x, _v), REF_VARIANT)

namespace REF_NAME(glsl_ref) {

#include "glsl_compat.hpp"

/* per-thread state of the fragment being shaded */
struct tex_state {
    const oracle_frame *frame;
    int nearest[5];            /* min/mag decision per texture unit for this fragment */
    int sampled[5];            /* did the fragment sample unit k, and with which direction */
    vec3 dir[5];
    int ncalls;
    vec3 linear;               /* argument of the final pow(color, 1/2.2) = linear colour */
};
static thread_local tex_state T;

static vec4 ref_texture_cube(int slot, vec3 d)
{
    float in[3] = { d.x, d.y, d.z }, rgb[3];
    T.sampled[slot] = 1; T.dir[slot] = d; T.ncalls++;
    oracle_texture_cube(T.frame, slot, in, T.nearest[slot], rgb);
    return vec4(rgb[0], rgb[1], rgb[2], 1.0f);
}
/* main() ends with `pow(color, vec3(1.0 / 2.2))` (fragment.shd:959): remember its argument,
 * that is the linear-RGB colour the parity bar (PSNR) is stated on */
static inline vec3 ref_pow_capture(vec3 a, vec3 b) { T.linear = a; return pow(a, b); }

#define uniform static
#define out static thread_local
#define main shader_main
#define pow(a, b) ref_pow_dispatch(a, b)
static inline float ref_pow_dispatch(float a, float b) { return (pow)(a, b); }
static inline vec3 ref_pow_dispatch(vec3 a, vec3 b) { return ref_pow_capture(a, b); }
#include "fragment.gen.inc"
#undef pow
#undef main
#undef out
#undef uniform

static void set_uniforms(const oracle_frame *f)
{
    in_time = f->time;
    in_screen_wdh = (float)f->fb_w;
    in_screen_hgt = (float)f->fb_h;
    env_reflection.slot = ORACLE_ENV_REFLECTION;        /* texture units, ShaderRendering.hs:164-176 */
    env_cos_1.slot = ORACLE_ENV_COS_1; env_cos_8.slot = ORACLE_ENV_COS_8;
    env_cos_64.slot = ORACLE_ENV_COS_64; env_cos_512.slot = ORACLE_ENV_COS_512;
    cornell_geom.texels = f->cornell;
}

struct frag_out { vec4 color; vec3 linear; int ncalls; int sampled[5]; vec3 dir[5]; };

static void shade(const oracle_frame *f, int px, int py, const int nearest[5], frag_out *o)
{
    T.frame = f; T.ncalls = 0;
    for (int k = 0; k < 5; k++) { T.nearest[k] = nearest[k]; T.sampled[k] = 0; }
    gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
    shader_main();
    o->color = frag_color; o->linear = T.linear; o->ncalls = T.ncalls;
    for (int k = 0; k < 5; k++) { o->sampled[k] = T.sampled[k]; o->dir[k] = T.dir[k]; }
}

} /* namespace */

using namespace REF_NAME(glsl_ref);

extern "C" {

/* fragment.shd main() for every pixel of [x0,x1)×[y0,y1), as oracle_render() */
int REF_NAME(ref_render)(const oracle_frame *f, int x0, int y0, int x1, int y1,
                         uint32_t *fb, float *linear_rgb, uint8_t *hit_mask, uint64_t *nearest_quads)
{
    if (!f || f->fb_w <= 0 || f->fb_h <= 0) return 1;
    if (x0 < 0 || y0 < 0 || x1 > f->fb_w || y1 > f->fb_h || x0 >= x1 || y0 >= y1) return 2;
    set_uniforms(f);
    const int qx0 = x0 >> 1, qx1 = (x1 + 1) >> 1, qy0 = y0 >> 1, qy1 = (y1 + 1) >> 1;
    uint64_t nq = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nq) num_threads(f->threads > 0 ? f->threads : omp_get_max_threads())
#endif
    for (int qy = qy0; qy < qy1; qy++) {
        for (int qx = qx0; qx < qx1; qx++) {
            frag_out o[4];
            int px[4], py[4], inside[4], near[5] = { 0, 0, 0, 0, 0 };
            const int linear_all[5] = { 0, 0, 0, 0, 0 };
            bool uniform_flow = true;
            for (int k = 0; k < 4; k++) {
                px[k] = qx * 2 + (k & 1); py[k] = qy * 2 + (k >> 1);
                inside[k] = px[k] < f->fb_w && py[k] < f->fb_h;
                if (inside[k]) shade(f, px[k], py[k], linear_all, &o[k]); else o[k] = o[0];
                for (int s = 0; s < 5; s++) if (o[k].sampled[s] != o[0].sampled[s]) uniform_flow = false;
            }
            /* C5: one min/mag decision per quad and texture unit when all four fragments took the
             * same branch (derivatives are undefined otherwise → LINEAR) */
            if (f->lod_policy == ORACLE_LOD_QUAD && inside[3] && uniform_flow) {
                bool any = false;
                for (int s = 0; s < 5; s++) {
                    if (!o[0].sampled[s] || f->env_face[s] <= 0) continue;
                    float d[12];
                    for (int k = 0; k < 4; k++) { d[k * 3] = o[k].dir[s].x; d[k * 3 + 1] = o[k].dir[s].y; d[k * 3 + 2] = o[k].dir[s].z; }
                    near[s] = oracle_quad_is_minified(f->env_face[s], d);
                    any = any || near[s];
                }
                if (any) {
                    nq++;
                    for (int k = 0; k < 4; k++) shade(f, px[k], py[k], near, &o[k]);
                }
            }
            for (int k = 0; k < 4; k++) {
                if (!inside[k] || px[k] < x0 || px[k] >= x1 || py[k] < y0 || py[k] >= y1) continue;
                size_t idx = (size_t)(py[k] - y0) * (size_t)(x1 - x0) + (size_t)(px[k] - x0);
                const vec4 c = o[k].color;
                if (fb) fb[idx] = rg_unorm8(c.x) | (rg_unorm8(c.y) << 8) | (rg_unorm8(c.z) << 16) | (rg_unorm8(c.w) << 24);
                if (linear_rgb) { linear_rgb[idx * 3] = o[k].linear.x; linear_rgb[idx * 3 + 1] = o[k].linear.y; linear_rgb[idx * 3 + 2] = o[k].linear.z; }
                if (hit_mask) hit_mask[idx] = (uint8_t)(o[k].ncalls != 1);   /* a miss samples env_reflection once (:823/:936) */
            }
        }
    }
    if (nearest_quads) *nearest_quads = nq;
    return 0;
}

/* ---- single functions of the shader, for lattice tests ------------------------------- */
static void set_scalar_uniforms(float time, float w, float h, const float *cornell)
{
    oracle_frame f; memset(&f, 0, sizeof f);
    f.time = time; f.fb_w = (int)w; f.fb_h = (int)h; f.cornell = cornell;
    set_uniforms(&f);
    in_screen_wdh = w; in_screen_hgt = h;
}
void REF_NAME(ref_distance_estimator)(float time, const float *cornell, const float *pts, float *out, int64_t n)
{
    set_scalar_uniforms(time, 1.0f, 1.0f, cornell);
    for (int64_t i = 0; i < n; i++) out[i] = distance_estimator(vec3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]));
}
/* out: n × (hit, t) */
void REF_NAME(ref_ray_march)(float time, const float *cornell, const float *origins, const float *dirs, float *out, int64_t n)
{
    set_scalar_uniforms(time, 1.0f, 1.0f, cornell);
    for (int64_t i = 0; i < n; i++) {
        float t = 0.0f, sg = 0.0f;
        bool hit = ray_march(vec3(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]), vec3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]), t, sg);
        out[i * 2] = hit ? 1.0f : 0.0f; out[i * 2 + 1] = hit ? t : 0.0f;
    }
}
void REF_NAME(ref_normal)(float time, const float *cornell, int central, const float *pts, float *out, int64_t n)
{
    set_scalar_uniforms(time, 1.0f, 1.0f, cornell);
    for (int64_t i = 0; i < n; i++) {
        vec3 p(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
        vec3 r = central ? normal_central_difference(p) : normal_backward_difference(p);
        out[i * 3] = r.x; out[i * 3 + 1] = r.y; out[i * 3 + 2] = r.z;
    }
}
void REF_NAME(ref_distance_ao)(float time, const float *cornell, const float *pts, const float *normals, float *out, int64_t n)
{
    set_scalar_uniforms(time, 1.0f, 1.0f, cornell);
    for (int64_t i = 0; i < n; i++)
        out[i] = distance_ao(vec3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]), vec3(normals[i * 3], normals[i * 3 + 1], normals[i * 3 + 2]));
}
void REF_NAME(ref_fresnel_conductor)(const float *cosi, float eta, float k, float *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) out[i] = fresnel_conductor(cosi[i], eta, k);
}
/* lookat (fragment.shd:829-838): cam[12] = x̂, ŷ, ẑ, eye */
void REF_NAME(ref_lookat)(const float eye[3], const float focus[3], const float up[3], float cam[12])
{
    mat4x4 m = lookat(vec3(eye[0], eye[1], eye[2]), vec3(focus[0], focus[1], focus[2]), vec3(up[0], up[1], up[2]));
    for (int c = 0; c < 4; c++) { cam[c * 3] = m.c[c].x; cam[c * 3 + 1] = m.c[c].y; cam[c * 3 + 2] = m.c[c].z; }
}
/* generate_ray (fragment.shd:840-871) for pixel (px, py) of a w×h frame; out = origin, dir */
void REF_NAME(ref_generate_ray)(const float cam[12], float w, float h, int px, int py, int ortho, float width_or_hfov, float out[6])
{
    set_scalar_uniforms(0.0f, w, h, 0);
    mat4x4 m(cam[0], cam[1], cam[2], 0.0f, cam[3], cam[4], cam[5], 0.0f, cam[6], cam[7], cam[8], 0.0f, cam[9], cam[10], cam[11], 1.0f);
    vec3 o, d;
    gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
    generate_ray(m, vec2(0.0f, 0.0f), ortho != 0, width_or_hfov, o, d);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = d.x; out[4] = d.y; out[5] = d.z;
}
/* out: n × (hit, tmin, tmax) */
void REF_NAME(ref_ray_sphere)(const float *origins, const float *dirs, float radius, float *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        float t0 = 0.0f, t1 = 0.0f;
        bool hit = ray_sphere(vec3(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]), vec3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]), vec3(0.0f, 0.0f, 0.0f), radius, t0, t1);
        out[i * 3] = hit ? 1.0f : 0.0f; out[i * 3 + 1] = hit ? t0 : 0.0f; out[i * 3 + 2] = hit ? t1 : 0.0f;
    }
}
void REF_NAME(ref_triplex_pow8)(const float *pts, float *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        vec3 r = triplex_pow8(vec3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]));
        out[i * 3] = r.x; out[i * 3 + 1] = r.y; out[i * 3 + 2] = r.z;
    }
}
void REF_NAME(ref_triplex_pow)(const float *pts, float power, float *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        vec3 r = triplex_pow(vec3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]), power);
        out[i * 3] = r.x; out[i * 3 + 1] = r.y; out[i * 3 + 2] = r.z;
    }
}
void REF_NAME(ref_de_triangle)(const float *pts, const float tri[9], float *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++)
        out[i] = de_triangle(vec3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]), vec3(tri[0], tri[1], tri[2]), vec3(tri[3], tri[4], tri[5]), vec3(tri[6], tri[7], tri[8]));
}
void REF_NAME(ref_smin)(const float *a, const float *b, float k, float *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) out[i] = smin(a[i], b[i], k);
}

} /* extern "C" */
