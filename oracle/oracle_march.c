/* oracle_march.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's per-pixel sphere tracer,
 * /root/reference/fragment.shd, one C function per GLSL function, each citing the
 * shader lines it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (librmdf_b200.so) never links, loads or calls it.
 *
 * PARITY STATUS: PINNED by oracle/_ref — the reference's own fragment.shd compiled by g++
 * (oracle/ref/, `make -C oracle ref`): every function below and whole frames per shader
 * variant are bit-identical to that compiled shader (tests/test_ref_shader_cpu.py), and
 * tests/golden/ref_shader.npz carries its outputs to the GPU box.  What stays a convention
 * is the platform around the shader (GLSL built-ins, GL sampler: C0-C6 below), which the
 * compiled shader shares through oracle/ref/glsl_compat.hpp.
 *
 * GLSL built-ins come from include/rmdf_glsl.h (deterministic IEEE-only
 * definitions); compile with -ffp-contract=off and without -ffast-math.
 *
 * Oracle conventions where GLSL/GL leave the result to the implementation
 * (all documented in DESIGN.md §"Oracle conventions"):
 *   C1  pow(r, power-1) with the compile-time constant power 8 (POWER8 variant)
 *       is r^7 by multiplication ((r²·r²)·r²)·r; every other pow is exp(y·log x).
 *   C2  normalize(v) = v·(1/sqrt(dot(v,v))); dot sums left to right.
 *   C3  mat3(camera)·v = (x̂·v.x + ŷ·v.y) + ẑ·v.z per component.
 *   C4  cube-map sampling: face by largest |component| (ties X > Y > Z); LINEAR =
 *       FP32 weights; seamless edges by re-selecting the face of the virtual
 *       texel centre on the extended face plane; RGB16F texels.
 *   C5  min = NEAREST / mag = LINEAR switch (GLHelpers.hs:105-106 TFMagOnly): one
 *       decision per 2×2 fragment quad from coarse derivatives of the
 *       face-projected coordinates (λ > 0 ⇔ ρ² > 1 ⇒ NEAREST).  Every texture()
 *       call of the shader sits inside non-uniform control flow (`if (hit)`,
 *       fragment.shd:740-826) where GLSL leaves derivatives undefined; the
 *       oracle applies the rule when all four fragments of the quad took the
 *       same branch and uses LINEAR otherwise.  lod_policy = 1 forces LINEAR.
 *   C6  float → RGBA8: clamp, round to nearest; NaN → 0.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/rmdf_glsl.h"
#include "oracle.h"

typedef rg_vec3 vec3;

/* per-thread work counters (algorithmic) */
typedef struct {
    uint64_t samples, in_sphere, hits, march_de, shade_de, mb_iters, nan_pixels, nearest_quads;
} counters_t;

typedef struct {
    const oracle_frame *f;
    float power;          /* general Mandelbulb power, fragment.shd:116-119 */
    counters_t c;
} octx;

/* ---- fragment.shd:21-33  primitives ---------------------------------------- */
static float de_sphere(vec3 pos, float r) { return rg_length(pos) - r; }
static float de_torus(vec3 pos, float torus_size, float torus_r)
{
    float qx = rg_length2(pos.x, pos.y) - torus_size;
    float qy = pos.z;
    return rg_length2(qx, qy) - torus_r;
}
static float de_rounded_box(vec3 pos, vec3 box, float r)
{
    vec3 d = rg_v3(rg_max(rg_abs(pos.x) - box.x, 0.0f),
                   rg_max(rg_abs(pos.y) - box.y, 0.0f),
                   rg_max(rg_abs(pos.z) - box.z, 0.0f));
    return rg_length(d) - r;
}

/* ---- fragment.shd:42-72  general triplex power ----------------------------- */
static vec3 triplex_pow(vec3 w, float power)
{
    float r = rg_length(w);
    float theta = rg_acos(w.z / r);
    float phi = rg_atan2(w.y, w.x);
    float zr = rg_pow(r, power);
    float st, ct, sp, cp;
    theta = theta * power;
    phi = phi * power;
    st = rg_sin(theta); ct = rg_cos(theta); sp = rg_sin(phi); cp = rg_cos(phi);
    return rg_v3(zr * (st * cp), zr * (st * sp), zr * ct);
}

/* ---- fragment.shd:74-99  power-8 special case ------------------------------ */
static vec3 triplex_pow8(vec3 w)
{
    float x = w.x; float x2 = x * x; float x4 = x2 * x2;
    float y = w.y; float y2 = y * y; float y4 = y2 * y2;
    float z = w.z; float z2 = z * z; float z4 = z2 * z2;

    float k3 = y2 + x2;
    float k2 = rg_inversesqrt(k3 * k3 * k3 * k3 * k3 * k3 * k3);
    float k1 = y4 + z4 + x4 - 6.0f * z2 * x2 - 6.0f * y2 * z2 + 2.0f * x2 * y2;
    float k4 = y2 - z2 + x2;

    return rg_v3(-8.0f * z * k4 * (y4 * y4 - 28.0f * y4 * y2 * x2 + 70.0f * y4 * x4 - 28.0f * y2 * x2 * x4 + x4 * x4) * k1 * k2,
                 64.0f * y * z * x * (y2 - x2) * k4 * (y4 - 6.0f * y2 * x2 + x4) * k1 * k2,
                 -16.0f * z2 * k3 * k4 * k4 + k1 * k1);
}

/* ---- fragment.shd:101-158 --------------------------------------------------- */
static float de_mandelbulb(octx *o, vec3 pos, int power8)
{
    const float bailout = 4.0f;
    const int iterations = 25;
    float power = power8 ? 8.0f : o->power;
    vec3 w;
    float dr = 1.0f, r = 0.0f;
    int i;

    pos = rg_v3(pos.z, pos.x, pos.y);          /* pos = pos.zxy  (:125) */
    w = pos;
    for (i = 0; i < iterations; i++) {
        r = rg_length(w);
        if (r > bailout) break;
        if (power8) w = triplex_pow8(w);
        else        w = triplex_pow(w, power);
        w = rg_add(w, pos);
        if (power8) {                          /* convention C1 */
            float r2 = r * r, r4 = r2 * r2;
            dr = ((r4 * r2) * r) * power * dr + 1.0f;
        } else {
            dr = rg_pow(r, power - 1.0f) * power * dr + 1.0f;
        }
        o->c.mb_iters++;
    }
    return 0.5f * rg_log(r) * r / dr;
}

/* ---- fragment.shd:312-372  point–triangle distance ------------------------- */
static float line_seg_min_dist_sq(vec3 a, vec3 b, vec3 p)
{
    vec3 ab = rg_sub(b, a);
    float len_sq = rg_dot(ab, ab);
    float t = rg_dot(rg_sub(p, a), ab) / len_sq;
    vec3 proj, d;
    t = rg_clamp(t, 0.0f, 1.0f);
    proj = rg_add(a, rg_scale(ab, t));
    d = rg_sub(p, proj);
    return rg_dot(d, d);
}
static int compute_barycentric(vec3 pos, vec3 v0, vec3 v1, vec3 v2, float *u, float *v)
{
    vec3 e0 = rg_sub(v2, v0), e1 = rg_sub(v1, v0), e2 = rg_sub(pos, v0);
    float dot00 = rg_dot(e0, e0), dot01 = rg_dot(e0, e1), dot02 = rg_dot(e0, e2);
    float dot11 = rg_dot(e1, e1), dot12 = rg_dot(e1, e2);
    float inv_denom = 1.0f / (dot00 * dot11 - dot01 * dot01);
    *u = (dot11 * dot02 - dot01 * dot12) * inv_denom;
    *v = (dot00 * dot12 - dot01 * dot02) * inv_denom;
    return (*u >= 0.0f) && (*v >= 0.0f) && (*u + *v < 1.0f);
}
static float de_triangle(vec3 pos, vec3 v0, vec3 v1, vec3 v2)
{
    float u, v;
    if (compute_barycentric(pos, v0, v1, v2, &u, &v)) {
        vec3 pp = rg_add(rg_add(rg_scale(v2, u), rg_scale(v1, v)), rg_scale(v0, 1.0f - (u + v)));
        return rg_distance(pos, pp);
    }
    return rg_sqrt(rg_min(line_seg_min_dist_sq(v0, v1, pos),
                   rg_min(line_seg_min_dist_sq(v0, v2, pos),
                          line_seg_min_dist_sq(v1, v2, pos))));
}
/* ---- fragment.shd:374-411 ---------------------------------------------------- */
static float de_cornell_box(octx *o, vec3 pos)
{
    const float *g = o->f->cornell;      /* texelFetch(cornell_geom, i*3+k, 0).xyz */
    float dist = 999.0f;
    int i;
    for (i = 0; i < 32; i++) {
        const float *t = g + i * 9;
        dist = rg_min(dist, de_triangle(pos, rg_v3(t[0], t[1], t[2]),
                                             rg_v3(t[3], t[4], t[5]),
                                             rg_v3(t[6], t[7], t[8])));
    }
    return dist;
}

/* ---- fragment.shd:413-418 ---------------------------------------------------- */
static float smin(float a, float b, float k)
{
    float res = rg_exp(-k * a) + rg_exp(-k * b);
    return -rg_log(res) / k;
}

/* ---- fragment.shd:420-458 ---------------------------------------------------- */
static float distance_estimator(octx *o, vec3 pos)
{
    switch (o->f->scene) {
    case ORACLE_SCENE_MB_POWER8:  return de_mandelbulb(o, pos, 1);
    case ORACLE_SCENE_MB_GENERAL: return de_mandelbulb(o, pos, 0);
    case ORACLE_SCENE_CORNELL:    return de_cornell_box(o, pos);
    default: {
        vec3 pzxy = rg_v3(pos.z, pos.x, pos.y), pyzx = rg_v3(pos.y, pos.z, pos.x);
        float d_sphere = de_sphere(pos, 0.4f);
        float d_torus = smin(smin(de_torus(pos, 0.85f, 0.1f),
                                  de_torus(pzxy, 0.85f, 0.1f), 64.0f),
                             de_torus(pyzx, 0.85f, 0.1f), 64.0f);
        float d_box = smin(smin(de_rounded_box(pos, rg_v3(0.8f, 0.06f, 0.06f), 0.03f),
                                de_rounded_box(pos, rg_v3(0.06f, 0.8f, 0.06f), 0.03f), 64.0f),
                           de_rounded_box(pos, rg_v3(0.06f, 0.06f, 0.8f), 0.03f), 64.0f);
        return smin(d_box, rg_min(d_sphere, d_torus), 64.0f);
    }
    }
}

/* ---- fragment.shd:463-470 ---------------------------------------------------- */
static vec3 normal_backward_difference(octx *o, vec3 pos)
{
    const float eps = 0.00001f;
    float c = distance_estimator(o, pos);
    float dx = c - distance_estimator(o, rg_sub(pos, rg_v3(eps, 0.0f, 0.0f)));
    float dy = c - distance_estimator(o, rg_sub(pos, rg_v3(0.0f, eps, 0.0f)));
    float dz = c - distance_estimator(o, rg_sub(pos, rg_v3(0.0f, 0.0f, eps)));
    o->c.shade_de += 4;
    return rg_normalize(rg_v3(dx, dy, dz));
}

/* ---- fragment.shd:471-480 (defined but unused by the reference: option) ------ */
static vec3 normal_central_difference(octx *o, vec3 pos)
{
    const float eps = 0.00001f;
    float dx = distance_estimator(o, rg_add(pos, rg_v3(eps, 0.0f, 0.0f))) - distance_estimator(o, rg_sub(pos, rg_v3(eps, 0.0f, 0.0f)));
    float dy = distance_estimator(o, rg_add(pos, rg_v3(0.0f, eps, 0.0f))) - distance_estimator(o, rg_sub(pos, rg_v3(0.0f, eps, 0.0f)));
    float dz = distance_estimator(o, rg_add(pos, rg_v3(0.0f, 0.0f, eps))) - distance_estimator(o, rg_sub(pos, rg_v3(0.0f, 0.0f, eps)));
    o->c.shade_de += 6;
    return rg_normalize(rg_v3(dx, dy, dz));
}

/* ---- fragment.shd:542-591 ---------------------------------------------------- */
static float ao_term(octx *o, vec3 p, vec3 n, float weight, float delta)
{
    float d = distance_estimator(o, rg_add(p, rg_scale(n, delta)));
    o->c.shade_de += 1;
    return weight * rg_clamp(1.0f - d / delta, 0.0f, 1.0f);
}
static float distance_ao(octx *o, vec3 p, vec3 n)
{
    float occl_sum = 0.0f;
    if (o->f->scene != ORACLE_SCENE_CORNELL) {
        occl_sum += ao_term(o, p, n, 0.5f, 0.016f);
        occl_sum += ao_term(o, p, n, 0.25f, 0.081f);
        occl_sum = 1.0f - occl_sum;
        occl_sum -= 0.29f;
        occl_sum *= 3.5f;
        occl_sum *= occl_sum;
        occl_sum = rg_clamp(occl_sum, 0.0f, 1.0f);
        return occl_sum;
    }
    occl_sum += ao_term(o, p, n, 0.1f, 0.1f);
    occl_sum += ao_term(o, p, n, 0.2f, 0.2f);
    occl_sum += ao_term(o, p, n, 0.125f, 0.4f);
    occl_sum += ao_term(o, p, n, 0.0625f, 0.5f);
    occl_sum = 1.0f - occl_sum;
    return occl_sum;
}

/* ---- fragment.shd:595-616 ---------------------------------------------------- */
static int ray_sphere(vec3 origin, vec3 dir, vec3 sphere_pos, float sphere_r, float *tmin, float *tmax)
{
    vec3 rs = rg_sub(sphere_pos, origin);
    float t = rg_dot(dir, rs);
    float a = rg_dot(rs, rs) - t * t;
    float r2 = sphere_r * sphere_r;
    float h;
    if (a > r2) return 0;
    h = rg_sqrt(r2 - a);
    *tmin = t - h;
    *tmax = t + h;
    return 1;
}

/* ---- fragment.shd:618-676 ---------------------------------------------------- */
static int ray_march(octx *o, vec3 origin, vec3 dir, float *t_out)
{
    const int MAX_STEPS = 128;
    const float MIN_DIST = 0.001f;
    float b_sphere_r, tspheremin, tspheremax, t;
    int steps;
    switch (o->f->scene) {
    case ORACLE_SCENE_MB_POWER8:  b_sphere_r = 1.15f; break;
    case ORACLE_SCENE_MB_GENERAL: b_sphere_r = 1.5f;  break;
    default:                      b_sphere_r = 1.0f;  break;
    }
    if (!ray_sphere(origin, dir, rg_v3(0.0f, 0.0f, 0.0f), b_sphere_r, &tspheremin, &tspheremax))
        return 0;
    o->c.in_sphere++;
    t = tspheremin;
    t = rg_max(0.0f, t);
    for (steps = 0; steps < MAX_STEPS; steps++) {
        vec3 pos = rg_add(origin, rg_scale(dir, t));
        float dist = distance_estimator(o, pos);
        o->c.march_de++;
        t += dist;
        if (t > tspheremax) return 0;
        if (dist < MIN_DIST) { *t_out = t; return 1; }
    }
    return 0;
}

/* ---- fragment.shd:694-724 ---------------------------------------------------- */
static float fresnel_conductor(float cosi, float eta, float k)
{
    float tmp = (eta * eta + k * k) * cosi * cosi;
    float r_parallel_2 = (tmp - (2.0f * eta * cosi) + 1.0f) /
                         (tmp + (2.0f * eta * cosi) + 1.0f);
    float tmp_f = eta * eta + k * k;
    float r_perpend_2 = (tmp_f - (2.0f * eta * cosi) + cosi * cosi) /
                        (tmp_f + (2.0f * eta * cosi) + cosi * cosi);
    return (r_parallel_2 + r_perpend_2) / 2.0f;
}
static float normalize_phong_lobe(float power) { return (power + 2.0f) / 2.0f; }

/* ---- samplerCube: OpenGL 3.3 §3.8.10 + conventions C4/C5 --------------------- */
/* face order +X,-X,+Y,-Y,+Z,-Z = HDREnvMap.hs:131-136 */
static int cube_face(vec3 d)
{
    float ax = rg_abs(d.x), ay = rg_abs(d.y), az = rg_abs(d.z);
    if (ax >= ay && ax >= az) return d.x < 0.0f ? 1 : 0;
    if (ay >= az)             return d.y < 0.0f ? 3 : 2;
    return d.z < 0.0f ? 5 : 4;
}
/* (sc, tc, ma) for a given face — GL spec table 3.19; ma keeps its sign */
static void cube_sctcma(int face, vec3 d, float *sc, float *tc, float *ma)
{
    switch (face) {
    case 0:  *sc = -d.z; *tc = -d.y; *ma = d.x;  break;
    case 1:  *sc =  d.z; *tc = -d.y; *ma = -d.x; break;
    case 2:  *sc =  d.x; *tc =  d.z; *ma = d.y;  break;
    case 3:  *sc =  d.x; *tc = -d.z; *ma = -d.y; break;
    case 4:  *sc =  d.x; *tc = -d.y; *ma = d.z;  break;
    default: *sc = -d.x; *tc = -d.y; *ma = -d.z; break;
    }
}
/* inverse of the above for ma = 1: direction of face-plane point (sc, tc) */
static vec3 cube_dir(int face, float sc, float tc)
{
    switch (face) {
    case 0:  return rg_v3(1.0f, -tc, -sc);
    case 1:  return rg_v3(-1.0f, -tc, sc);
    case 2:  return rg_v3(sc, 1.0f, tc);
    case 3:  return rg_v3(sc, -1.0f, -tc);
    case 4:  return rg_v3(sc, -tc, 1.0f);
    default: return rg_v3(-sc, -tc, -1.0f);
    }
}
static vec3 cube_texel(const float *env, int w, int face, int i, int j)
{
    const float *p;
    if (i < 0 || j < 0 || i >= w || j >= w) {
        /* seamless: virtual texel centre on the extended plane → its real face */
        float fw = (float)w;
        float sc = ((float)i + 0.5f) / fw * 2.0f - 1.0f;
        float tc = ((float)j + 0.5f) / fw * 2.0f - 1.0f;
        vec3 d = cube_dir(face, sc, tc);
        float s2, t2, ma;
        face = cube_face(d);
        cube_sctcma(face, d, &s2, &t2, &ma);
        i = (int)rg_floor(0.5f * (s2 / ma + 1.0f) * fw);
        j = (int)rg_floor(0.5f * (t2 / ma + 1.0f) * fw);
        if (i < 0) i = 0; if (i > w - 1) i = w - 1;
        if (j < 0) j = 0; if (j > w - 1) j = w - 1;
    }
    p = env + ((size_t)(face * w + j) * w + i) * 3;
    return rg_v3(p[0], p[1], p[2]);
}
static vec3 texture_cube(const oracle_frame *f, int slot, vec3 d, int nearest)
{
    int w = f->env_face[slot];
    const float *env = f->env[slot];
    int face;
    float sc, tc, ma, fw, u, v;
    if (w <= 0 || !env) return rg_v3(1.0f, 1.0f, 1.0f);   /* "no env lighting": constant 1 (D6) */
    face = cube_face(d);
    cube_sctcma(face, d, &sc, &tc, &ma);
    fw = (float)w;
    u = 0.5f * (sc / ma + 1.0f) * fw;
    v = 0.5f * (tc / ma + 1.0f) * fw;
    if (!(u == u)) u = 0.0f;                 /* NaN direction → texel coordinate 0 */
    if (!(v == v)) v = 0.0f;
    if (nearest) {
        int i = (int)rg_floor(u), j = (int)rg_floor(v);
        if (i < 0) i = 0; if (i > w - 1) i = w - 1;
        if (j < 0) j = 0; if (j > w - 1) j = w - 1;
        return cube_texel(env, w, face, i, j);
    } else {
        float fu = u - 0.5f, fv = v - 0.5f;
        float fi = rg_floor(fu), fj = rg_floor(fv);
        float a = fu - fi, b = fv - fj;
        int i0 = (int)fi, j0 = (int)fj;
        vec3 t00 = cube_texel(env, w, face, i0, j0);
        vec3 t10 = cube_texel(env, w, face, i0 + 1, j0);
        vec3 t01 = cube_texel(env, w, face, i0, j0 + 1);
        vec3 t11 = cube_texel(env, w, face, i0 + 1, j0 + 1);
        float ia = 1.0f - a, ib = 1.0f - b;
        return rg_v3((t00.x * ia + t10.x * a) * ib + (t01.x * ia + t11.x * a) * b,
                     (t00.y * ia + t10.y * a) * ib + (t01.y * ia + t11.y * a) * b,
                     (t00.z * ia + t10.z * a) * ib + (t01.z * ia + t11.z * a) * b);
    }
}
/* convention C5: one min/mag decision for the quad; d[0..3] = (0,0),(1,0),(0,1),(1,1) */
static int quad_is_minified(int w, const vec3 d[4])
{
    int face = cube_face(d[0]), k;
    float fw = (float)w, u[3], v[3];
    float dudx, dvdx, dudy, dvdy, rx, ry;
    for (k = 0; k < 3; k++) {
        float sc, tc, ma;
        cube_sctcma(face, d[k], &sc, &tc, &ma);
        if (!(ma > 0.0f)) return 1;               /* left the face's hemisphere */
        u[k] = 0.5f * (sc / ma + 1.0f) * fw;
        v[k] = 0.5f * (tc / ma + 1.0f) * fw;
    }
    dudx = u[1] - u[0]; dvdx = v[1] - v[0];
    dudy = u[2] - u[0]; dvdy = v[2] - v[0];
    rx = dudx * dudx + dvdx * dvdx;
    ry = dudy * dudy + dvdy * dvdy;
    return rg_max(rx, ry) > 1.0f;
}

/* ---- camera: fragment.shd:829-838, 884-902 ----------------------------------- */
void oracle_camera(int scene, float in_time, float cam[12])
{
    vec3 cam_pos, zaxis, xaxis, yaxis, up = rg_v3(0.0f, 1.0f, 0.0f);
    if (scene == ORACLE_SCENE_CORNELL) {
        cam_pos.x = rg_sin(in_time / 2.0f) * 0.4f;
        cam_pos.z = -2.0f;
        cam_pos.y = rg_cos(in_time / 2.0f) * 0.4f;
    } else {
        cam_pos.x = rg_sin(in_time / 3.0f);
        cam_pos.z = rg_cos(in_time / 3.0f);
        cam_pos.y = rg_cos(in_time / 4.0f);
        cam_pos = rg_scale(rg_normalize(cam_pos), 2.414213562373095f);
    }
    zaxis = rg_normalize(rg_sub(cam_pos, rg_v3(0.0f, 0.0f, 0.0f)));
    xaxis = rg_normalize(rg_cross(up, zaxis));
    yaxis = rg_cross(zaxis, xaxis);
    cam[0] = xaxis.x; cam[1] = xaxis.y; cam[2] = xaxis.z;
    cam[3] = yaxis.x; cam[4] = yaxis.y; cam[5] = yaxis.z;
    cam[6] = zaxis.x; cam[7] = zaxis.y; cam[8] = zaxis.z;
    cam[9] = cam_pos.x; cam[10] = cam_pos.y; cam[11] = cam_pos.z;
}
float oracle_mb_power(float in_time)            /* fragment.shd:116-119 */
{
    float pow_offs = rg_mod(in_time / 2.0f, 9.0f);
    if (pow_offs > 4.5f) pow_offs = 9.0f - pow_offs;
    return pow_offs + 2.0f;
}

/* ---- fragment.shd:840-871; the reference builds the perspective branch (CAM_ORTHO off, :906) -- */
static void generate_ray(const float cam[12], float w, float h, int px, int py, int ortho, vec3 *origin, vec3 *dir)
{
    float ndc_x = ((float)px + 0.5f + 0.0f) / w * 2.0f - 1.0f;   /* gl_FragCoord = px + 0.5 */
    float ndc_y = ((float)py + 0.5f + 0.0f) / h * 2.0f - 1.0f;
    float aspect = w / h;
    if (ortho) {                                                 /* generate_ray(camera, 0, true, 2.0), :908 */
        float width = 2.0f, height = width / aspect;
        float a = ndc_x * (width / 2.0f), b = ndc_y * (height / 2.0f);
        /* (camera * vec4(a, b, 0, 1)).xyz, columns x̂ ŷ ẑ eye summed left to right */
        origin->x = ((cam[0] * a + cam[3] * b) + cam[6] * 0.0f) + cam[9] * 1.0f;
        origin->y = ((cam[1] * a + cam[4] * b) + cam[7] * 0.0f) + cam[10] * 1.0f;
        origin->z = ((cam[2] * a + cam[5] * b) + cam[8] * 0.0f) + cam[11] * 1.0f;
        dir->x = (cam[0] * 0.0f + cam[3] * 0.0f) + cam[6] * -1.0f;   /* mat3(camera) * vec3(0,0,-1) */
        dir->y = (cam[1] * 0.0f + cam[4] * 0.0f) + cam[7] * -1.0f;
        dir->z = (cam[2] * 0.0f + cam[5] * 0.0f) + cam[8] * -1.0f;
    } else {
        float hfov = rg_radians(45.0f * 1.5f);
        float fov_xs = rg_tan(hfov / 2.0f);
        vec3 v = rg_normalize(rg_v3(ndc_x * fov_xs, ndc_y * fov_xs / aspect, -1.0f));
        *origin = rg_v3(cam[9], cam[10], cam[11]);
        dir->x = (cam[0] * v.x + cam[3] * v.y) + cam[6] * v.z;       /* convention C3 */
        dir->y = (cam[1] * v.x + cam[4] * v.y) + cam[7] * v.z;
        dir->z = (cam[2] * v.x + cam[5] * v.y) + cam[8] * v.z;
    }
}

/* ---- fragment.shd:253-310 Möller–Trumbore, for RAY_TRACING_TEST --------------------- */
static int intersect_triangle(vec3 orig, vec3 dir, vec3 vert0, vec3 vert1, vec3 vert2, float *t, float *u, float *v)
{
    const float JGT_RAYTRI_EPSILON = 0.000001f;
    vec3 edge1 = rg_sub(vert1, vert0), edge2 = rg_sub(vert2, vert0);
    vec3 pvec = rg_cross(dir, edge2), tvec, qvec;
    float det = rg_dot(edge1, pvec), inv_det;
    if (det > -JGT_RAYTRI_EPSILON && det < JGT_RAYTRI_EPSILON) return 0;
    inv_det = 1.0f / det;
    tvec = rg_sub(orig, vert0);
    *u = rg_dot(tvec, pvec) * inv_det;
    if (*u < 0.0f || *u > 1.0f) return 0;
    qvec = rg_cross(tvec, edge1);
    *v = rg_dot(dir, qvec) * inv_det;
    if (*v < 0.0f || *u + *v > 1.0f) return 0;
    *t = rg_dot(edge2, qvec) * inv_det;
    return 1;
}

/* ---- fragment.shd:726-827 render_ray, split at the texture() calls so that the
 *      quad-level min/mag decision (C5) can be taken between the two halves ---- */
typedef struct {
    int hit;              /* ray_march result (RAY_TRACING_TEST: a triangle was hit) */
    int rt;               /* RAY_TRACING_TEST fragment: colour is (n+1)/2, no texture fetch on a hit */
    vec3 dir, n, refl;
    float ao, fresnel;
} surf_t;

static void trace_sample(octx *o, const float cam[12], int px, int py, surf_t *s)
{
    vec3 origin;
    float t = 0.0f;
    generate_ray(cam, (float)o->f->fb_w, (float)o->f->fb_h, px, py, o->f->flags & ORACLE_FLAG_CAM_ORTHO, &origin, &s->dir);
    o->c.samples++;
    s->rt = 0;
    if (o->f->flags & ORACLE_FLAG_RAY_TRACING_TEST) {            /* fragment.shd:915-937 */
        const float *g = o->f->cornell;
        float mint = 999.0f;
        int i;
        s->rt = 1; s->hit = 0; s->n = rg_v3(0.0f, 0.0f, 0.0f);
        for (i = 0; i < 32; i++) {
            vec3 v0 = rg_v3(g[i * 9], g[i * 9 + 1], g[i * 9 + 2]), v1 = rg_v3(g[i * 9 + 3], g[i * 9 + 4], g[i * 9 + 5]);
            vec3 v2 = rg_v3(g[i * 9 + 6], g[i * 9 + 7], g[i * 9 + 8]);
            float tt, u, v;
            if (intersect_triangle(origin, s->dir, v0, v1, v2, &tt, &u, &v))
                if (tt < mint) { s->n = rg_normalize(rg_cross(rg_sub(v1, v0), rg_sub(v2, v0))); mint = tt; }
        }
        s->hit = !(mint == 999.0f);
        if (s->hit) o->c.hits++;
        return;
    }
    s->hit = ray_march(o, origin, s->dir, &t);
    if (s->hit) {
        vec3 isec_pos = rg_add(origin, rg_scale(s->dir, t));
        vec3 np = rg_sub(isec_pos, rg_scale(s->dir, 0.00001f));
        o->c.hits++;
        s->n = (o->f->flags & ORACLE_FLAG_NORMAL_CENTRAL) ? normal_central_difference(o, np) : normal_backward_difference(o, np);
        s->ao = distance_ao(o, isec_pos, s->n);
        s->fresnel = fresnel_conductor(rg_dot(rg_neg(s->dir), s->n), 0.4f, 0.8f);
        s->refl = rg_reflect(s->dir, s->n);
    }
}
static vec3 shade_sample(const oracle_frame *f, const surf_t *s, int near_cos1, int near_cos8, int near_refl)
{
    if (s->rt && s->hit)                                          /* (n + 1) * 0.5, :936 */
        return rg_v3((s->n.x + 1.0f) * 0.5f, (s->n.y + 1.0f) * 0.5f, (s->n.z + 1.0f) * 0.5f);
    if (s->hit) {
        const float diff_weight = 0.5f;
        const vec3 diff_col = { 1.0f, 0.8f, 0.8f }, spec_col = { 0.8f, 0.8f, 1.0f };
        const float spec_weight = 1.0f - diff_weight;
        vec3 e1 = texture_cube(f, ORACLE_ENV_COS_1, s->n, near_cos1);
        vec3 e8 = texture_cube(f, ORACLE_ENV_COS_8, s->refl, near_cos8);
        vec3 er = texture_cube(f, ORACLE_ENV_REFLECTION, s->refl, near_refl);
        float npl = normalize_phong_lobe(8.0f);
        vec3 t1 = rg_scale(rg_mul(e1, diff_col), diff_weight);
        vec3 t2 = rg_scale(rg_scale(rg_scale(rg_mul(e8, spec_col), npl), s->fresnel), spec_weight);
        vec3 t3 = rg_scale(rg_scale(rg_scale(er, spec_weight), s->fresnel), 0.1f);
        return rg_scale(rg_scale(rg_add(rg_add(t1, t2), t3), 3.0f), s->ao);
    }
    return texture_cube(f, ORACLE_ENV_REFLECTION, s->dir, near_refl);   /* BG_GRADIENT, :823 */
}

/* ---- fragment.shd:873-966 main, over a rectangle of the frame buffer --------- */
int oracle_render(const oracle_frame *f, int x0, int y0, int x1, int y1,
                  uint32_t *fb, float *linear_rgb, uint8_t *hit_mask, oracle_counters *out)
{
    float cam[12];
    int qy0, qy1, qx0, qx1, qy;
    counters_t total;
    float power;
    if (!f || f->fb_w <= 0 || f->fb_h <= 0 || f->scene < 0 || f->scene > 3) return 1;
    if (x0 < 0 || y0 < 0 || x1 > f->fb_w || y1 > f->fb_h || x0 >= x1 || y0 >= y1) return 2;
    if ((f->scene == ORACLE_SCENE_CORNELL || (f->flags & ORACLE_FLAG_RAY_TRACING_TEST)) && !f->cornell) return 3;
    oracle_camera(f->scene, f->time, cam);
    power = oracle_mb_power(f->time);
    memset(&total, 0, sizeof total);
    qx0 = x0 >> 1; qx1 = (x1 + 1) >> 1; qy0 = y0 >> 1; qy1 = (y1 + 1) >> 1;

#ifdef _OPENMP
#pragma omp parallel num_threads(f->threads > 0 ? f->threads : omp_get_max_threads())
#endif
    {
        octx o;
        int qx, k;
        memset(&o, 0, sizeof o);
        o.f = f; o.power = power;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (qy = qy0; qy < qy1; qy++) {
            for (qx = qx0; qx < qx1; qx++) {
                surf_t s[4];
                int px[4], py[4], inside[4];
                int all_hit = 1, all_miss = 1;
                int near_c1 = 0, near_c8 = 0, near_rf = 0;
                /* the GL rasteriser shades whole 2×2 quads; pixels of the quad that
                 * fall outside the frame buffer do not exist (odd dimensions) */
                for (k = 0; k < 4; k++) {
                    px[k] = qx * 2 + (k & 1); py[k] = qy * 2 + (k >> 1);
                    inside[k] = px[k] < f->fb_w && py[k] < f->fb_h;
                    if (inside[k]) trace_sample(&o, cam, px[k], py[k], &s[k]);
                    else { s[k] = s[0]; }
                    if (s[k].hit) all_miss = 0; else all_hit = 0;
                }
                if (f->lod_policy == ORACLE_LOD_QUAD && inside[3] && (all_hit || all_miss)) {
                    vec3 d[4];
                    if (all_hit && s[0].rt) {
                        /* ray-tracing test: hits do not sample a texture */
                    } else if (all_hit) {
                        for (k = 0; k < 4; k++) d[k] = s[k].n;
                        if (f->env_face[ORACLE_ENV_COS_1] > 0) near_c1 = quad_is_minified(f->env_face[ORACLE_ENV_COS_1], d);
                        for (k = 0; k < 4; k++) d[k] = s[k].refl;
                        if (f->env_face[ORACLE_ENV_COS_8] > 0) near_c8 = quad_is_minified(f->env_face[ORACLE_ENV_COS_8], d);
                        if (f->env_face[ORACLE_ENV_REFLECTION] > 0) near_rf = quad_is_minified(f->env_face[ORACLE_ENV_REFLECTION], d);
                    } else {
                        for (k = 0; k < 4; k++) d[k] = s[k].dir;
                        if (f->env_face[ORACLE_ENV_REFLECTION] > 0) near_rf = quad_is_minified(f->env_face[ORACLE_ENV_REFLECTION], d);
                    }
                    if (near_c1 || near_c8 || near_rf) o.c.nearest_quads++;
                }
                for (k = 0; k < 4; k++) {
                    vec3 color, gamma;
                    size_t idx;
                    if (!inside[k] || px[k] < x0 || px[k] >= x1 || py[k] < y0 || py[k] >= y1) continue;
                    color = shade_sample(f, &s[k], near_c1, near_c8, near_rf);
                    /* GAMMA_CORRECT, fragment.shd:956-960; alpha = 1 */
                    gamma = rg_v3(rg_pow(color.x, 1.0f / 2.2f), rg_pow(color.y, 1.0f / 2.2f), rg_pow(color.z, 1.0f / 2.2f));
                    if (color.x != color.x || color.y != color.y || color.z != color.z) o.c.nan_pixels++;
                    idx = (size_t)(py[k] - y0) * (size_t)(x1 - x0) + (size_t)(px[k] - x0);
                    if (fb) fb[idx] = rg_unorm8(gamma.x) | (rg_unorm8(gamma.y) << 8) | (rg_unorm8(gamma.z) << 16) | 0xff000000u;
                    if (linear_rgb) { linear_rgb[idx * 3] = color.x; linear_rgb[idx * 3 + 1] = color.y; linear_rgb[idx * 3 + 2] = color.z; }
                    if (hit_mask) hit_mask[idx] = (uint8_t)s[k].hit;
                }
            }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            total.samples += o.c.samples; total.in_sphere += o.c.in_sphere; total.hits += o.c.hits;
            total.march_de += o.c.march_de; total.shade_de += o.c.shade_de; total.mb_iters += o.c.mb_iters;
            total.nan_pixels += o.c.nan_pixels; total.nearest_quads += o.c.nearest_quads;
        }
    }
    if (out) {
        out->samples = total.samples; out->in_sphere = total.in_sphere; out->hits = total.hits;
        out->march_de = total.march_de; out->shade_de = total.shade_de; out->mb_iters = total.mb_iters;
        out->nan_pixels = total.nan_pixels; out->nearest_quads = total.nearest_quads;
    }
    return 0;
}

/* Single distance-estimator evaluation, for known-answer tests. */
float oracle_distance(int scene, float in_time, const float *cornell, float x, float y, float z)
{
    oracle_frame f;
    octx o;
    memset(&f, 0, sizeof f); memset(&o, 0, sizeof o);
    f.scene = scene; f.time = in_time; f.cornell = cornell;
    o.f = &f; o.power = oracle_mb_power(in_time);
    return distance_estimator(&o, rg_v3(x, y, z));
}

/* ---- single functions, for the lattice comparison with oracle/_ref (the reference's own
 *      shader text compiled by g++; tests/test_ref_shader_cpu.py) ------------------------- */
static void fn_ctx(octx *o, oracle_frame *f, int scene, float in_time, const float *cornell)
{
    memset(f, 0, sizeof *f); memset(o, 0, sizeof *o);
    f->scene = scene; f->time = in_time; f->cornell = cornell;
    o->f = f; o->power = oracle_mb_power(in_time);
}
void oracle_distance_n(int scene, float in_time, const float *cornell, const float *pts, float *out, int64_t n)
{
    oracle_frame f; octx o; int64_t i;
    fn_ctx(&o, &f, scene, in_time, cornell);
    for (i = 0; i < n; i++) out[i] = distance_estimator(&o, rg_v3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]));
}
void oracle_ray_march_n(int scene, float in_time, const float *cornell, const float *origins, const float *dirs, float *out, int64_t n)
{
    oracle_frame f; octx o; int64_t i;
    fn_ctx(&o, &f, scene, in_time, cornell);
    for (i = 0; i < n; i++) {
        float t = 0.0f;
        int hit = ray_march(&o, rg_v3(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]), rg_v3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]), &t);
        out[i * 2] = hit ? 1.0f : 0.0f; out[i * 2 + 1] = hit ? t : 0.0f;
    }
}
void oracle_normal_n(int scene, float in_time, const float *cornell, int central, const float *pts, float *out, int64_t n)
{
    oracle_frame f; octx o; int64_t i;
    fn_ctx(&o, &f, scene, in_time, cornell);
    for (i = 0; i < n; i++) {
        vec3 p = rg_v3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
        vec3 r = central ? normal_central_difference(&o, p) : normal_backward_difference(&o, p);
        out[i * 3] = r.x; out[i * 3 + 1] = r.y; out[i * 3 + 2] = r.z;
    }
}
void oracle_distance_ao_n(int scene, float in_time, const float *cornell, const float *pts, const float *normals, float *out, int64_t n)
{
    oracle_frame f; octx o; int64_t i;
    fn_ctx(&o, &f, scene, in_time, cornell);
    for (i = 0; i < n; i++)
        out[i] = distance_ao(&o, rg_v3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]), rg_v3(normals[i * 3], normals[i * 3 + 1], normals[i * 3 + 2]));
}
void oracle_fresnel_conductor_n(const float *cosi, float eta, float k, float *out, int64_t n)
{
    int64_t i;
    for (i = 0; i < n; i++) out[i] = fresnel_conductor(cosi[i], eta, k);
}
void oracle_generate_ray(const float cam[12], float w, float h, int px, int py, int ortho, float out[6])
{
    vec3 o, d;
    generate_ray(cam, w, h, px, py, ortho, &o, &d);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = d.x; out[4] = d.y; out[5] = d.z;
}
void oracle_ray_sphere_n(const float *origins, const float *dirs, float radius, float *out, int64_t n)
{
    int64_t i;
    for (i = 0; i < n; i++) {
        float t0 = 0.0f, t1 = 0.0f;
        int hit = ray_sphere(rg_v3(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]), rg_v3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]), rg_v3(0.0f, 0.0f, 0.0f), radius, &t0, &t1);
        out[i * 3] = hit ? 1.0f : 0.0f; out[i * 3 + 1] = hit ? t0 : 0.0f; out[i * 3 + 2] = hit ? t1 : 0.0f;
    }
}
void oracle_triplex_pow_n(const float *pts, float power, int power8, float *out, int64_t n)
{
    int64_t i;
    for (i = 0; i < n; i++) {
        vec3 w = rg_v3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
        vec3 r = power8 ? triplex_pow8(w) : triplex_pow(w, power);
        out[i * 3] = r.x; out[i * 3 + 1] = r.y; out[i * 3 + 2] = r.z;
    }
}
void oracle_de_triangle_n(const float *pts, const float tri[9], float *out, int64_t n)
{
    int64_t i;
    for (i = 0; i < n; i++)
        out[i] = de_triangle(rg_v3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]), rg_v3(tri[0], tri[1], tri[2]), rg_v3(tri[3], tri[4], tri[5]), rg_v3(tri[6], tri[7], tri[8]));
}
void oracle_smin_n(const float *a, const float *b, float k, float *out, int64_t n)
{
    int64_t i;
    for (i = 0; i < n; i++) out[i] = smin(a[i], b[i], k);
}
/* the GL cube sampler (conventions C4/C5), shared with the oracle/_ref harness: GL fixed
 * function is not part of the reference's source */
void oracle_texture_cube(const oracle_frame *f, int slot, const float d[3], int nearest, float rgb[3])
{
    vec3 c = texture_cube(f, slot, rg_v3(d[0], d[1], d[2]), nearest);
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
int oracle_quad_is_minified(int face_w, const float d[12])
{
    vec3 v[4]; int k;
    for (k = 0; k < 4; k++) v[k] = rg_v3(d[k * 3], d[k * 3 + 1], d[k * 3 + 2]);
    return quad_is_minified(face_w, v);
}

/* Element-wise access to the deterministic built-ins of include/rmdf_glsl.h, for
 * tests/test_glsl_builtins.py (accuracy against libm in double precision). */
void oracle_glsl_eval(int fn, const float *x, const float *y, float *out, int64_t n)
{
    int64_t i;
    for (i = 0; i < n; i++) {
        switch (fn) {
        case 0:  out[i] = rg_log(x[i]); break;
        case 1:  out[i] = rg_exp(x[i]); break;
        case 2:  out[i] = rg_pow(x[i], y[i]); break;
        case 3:  out[i] = rg_sin(x[i]); break;
        case 4:  out[i] = rg_cos(x[i]); break;
        case 5:  out[i] = rg_tan(x[i]); break;
        case 6:  out[i] = rg_acos(x[i]); break;
        case 7:  out[i] = rg_atan2(x[i], y[i]); break;
        case 8:  out[i] = rg_inversesqrt(x[i]); break;
        case 9:  out[i] = rg_round_to_half(x[i]); break;
        case 10: out[i] = (float)rg_unorm8(x[i]); break;
        case 11: out[i] = rg_mod(x[i], y[i]); break;
        default: out[i] = 0.0f; break;
        }
    }
}
