/* oracle.h — C interface of the CPU oracle (test infrastructure, see oracle_march.c).
 * Loaded with ctypes by oracle/oracle.py; never by the product. */
#ifndef RMDF_ORACLE_H
#define RMDF_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* FragmentShader enum order, ShaderRendering.hs:46-47 */
enum { ORACLE_SCENE_CORNELL = 0, ORACLE_SCENE_TEST = 1, ORACLE_SCENE_MB_POWER8 = 2, ORACLE_SCENE_MB_GENERAL = 3 };
/* texture-unit order of srEnvCubeMaps, ShaderRendering.hs:86-91,173-174 */
enum { ORACLE_ENV_REFLECTION = 0, ORACLE_ENV_COS_1 = 1, ORACLE_ENV_COS_8 = 2, ORACLE_ENV_COS_64 = 3, ORACLE_ENV_COS_512 = 4 };
enum { ORACLE_LOD_QUAD = 0, ORACLE_LOD_ALWAYS_LINEAR = 1 };
/* compile-time switches of the shader that are OFF in the reference */
enum { ORACLE_FLAG_NORMAL_CENTRAL = 1,   /* normal_central_difference, fragment.shd:471-480 (unused at :751) */
       ORACLE_FLAG_CAM_ORTHO = 2,        /* #define CAM_ORTHO, fragment.shd:906-908, 853-861 */
       ORACLE_FLAG_RAY_TRACING_TEST = 4  /* #define RAY_TRACING_TEST, fragment.shd:915-937 (Cornell geometry) */ };

typedef struct {
    int32_t scene;            /* ORACLE_SCENE_* */
    int32_t fb_w, fb_h;       /* in_screen_wdh / in_screen_hgt */
    float   time;             /* in_time */
    int32_t lod_policy;       /* ORACLE_LOD_* */
    int32_t threads;          /* 0 = all */
    int32_t flags;            /* ORACLE_FLAG_*; 0 = the reference's build */
    int32_t env_face[5];      /* cube face width per slot; 0 = constant 1.0 */
    const float *env[5];      /* 6·w·w·3 floats (RGB16F values widened), faces +X,-X,+Y,-Y,+Z,-Z */
    const float *cornell;     /* 96·3 floats (cornell_geom texels) */
} oracle_frame;

typedef struct {
    uint64_t samples, in_sphere, hits, march_de, shade_de, mb_iters, nan_pixels, nearest_quads;
} oracle_counters;

/* fragment.shd main() over the FB rectangle [x0,x1)×[y0,y1); row 0 = bottom.
 * fb: (x1-x0)·(y1-y0) words 0xAABBGGRR; linear_rgb: pre-gamma colour (3 floats per
 * sample); hit_mask: ray_march result.  Any of the three may be NULL. */
int   oracle_render(const oracle_frame *f, int x0, int y0, int x1, int y1,
                    uint32_t *fb, float *linear_rgb, uint8_t *hit_mask, oracle_counters *out);
void  oracle_camera(int scene, float in_time, float cam[12]);
float oracle_mb_power(float in_time);
float oracle_distance(int scene, float in_time, const float *cornell, float x, float y, float z);
void  oracle_glsl_eval(int fn, const float *x, const float *y, float *out, int64_t n);

/* single functions of fragment.shd over n inputs (lattice tests against oracle/_ref) */
void oracle_distance_n(int scene, float in_time, const float *cornell, const float *pts, float *out, int64_t n);
void oracle_ray_march_n(int scene, float in_time, const float *cornell, const float *origins, const float *dirs, float *out /*n·(hit,t)*/, int64_t n);
void oracle_normal_n(int scene, float in_time, const float *cornell, int central, const float *pts, float *out, int64_t n);
void oracle_distance_ao_n(int scene, float in_time, const float *cornell, const float *pts, const float *normals, float *out, int64_t n);
void oracle_fresnel_conductor_n(const float *cosi, float eta, float k, float *out, int64_t n);
void oracle_generate_ray(const float cam[12], float w, float h, int px, int py, int ortho, float out[6]);
void oracle_ray_sphere_n(const float *origins, const float *dirs, float radius, float *out /*n·(hit,tmin,tmax)*/, int64_t n);
void oracle_triplex_pow_n(const float *pts, float power, int power8, float *out, int64_t n);
void oracle_de_triangle_n(const float *pts, const float tri[9], float *out, int64_t n);
void oracle_smin_n(const float *a, const float *b, float k, float *out, int64_t n);
/* GL cube sampler (conventions C4/C5) */
void oracle_texture_cube(const oracle_frame *f, int slot, const float d[3], int nearest, float rgb[3]);
int  oracle_quad_is_minified(int face_w, const float d[12]);

/* HDREnvMap.hs / CoordTransf.hs / CornellBox.hs (oracle_envmap.c) */
int  oracle_hdr_decode(const uint8_t *file, int64_t nbytes, int *w, int *h, float *rgb, int64_t cap_floats);
int64_t oracle_hdr_encode(const float *rgb, int w, int h, uint8_t *out, int64_t cap);
void oracle_rgbe_roundtrip(const float *rgb, int64_t npix, float *out);
int  oracle_resize_height(int srcw, int srch, int dstw);
void oracle_resize(const float *src, int srcw, int srch, int dstw, float *dst);
void oracle_convolve(const float *src, int w, int h, float power, float *dst, int threads);
void oracle_convolve_rows(const float *src, int w, int h, float power, float *dst, int y0, int y1, int threads);
void oracle_convolve_sweep_rows(const float *src, int w, int h, const float *powers, int np, float *dst, int y0, int y1);
void oracle_latlong_to_cube(const float *latlong, int w, int h, float *faces /*6·fw·fw·3*/, int round_half);
void oracle_cube_pixel_to_dir(int face, int w, int x, int y, float dir[3]);
void oracle_test_latlong_envmap(float *rgb /*512·256·3*/);
void oracle_cornell_vertices(float *xyz /*96·3*/);
int  oracle_make_segments(int nseg, int low, int high, int *out_pairs, int cap_pairs);

/* Fractal2D.hs (oracle_julia.c) */
void oracle_julia(int w, int h, uint32_t *fb, int smooth, double tick, int y0, int y1, int threads);
void oracle_julia_segments(int w, int h, uint32_t *fb, int smooth, double tick, int y0, int y1, int nseg);

/* FrameBuffer.hs mip resolve + screenshot orientation (oracle_fb.c) */
void oracle_mip_resolve(const uint32_t *fb, int w, int h, int levels, uint32_t *out);
void oracle_screenshot_rows(const uint32_t *fb, int w, int h, uint8_t *rgba);

#ifdef __cplusplus
}
#endif
#endif
