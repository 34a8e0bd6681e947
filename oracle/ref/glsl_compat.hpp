/* glsl_compat.hpp — the GLSL 3.30 surface that /root/reference/fragment.shd uses, as C++.
 *
 * TEST INFRASTRUCTURE (oracle/_ref recipe).  With this header the reference's fragment
 * shader compiles with g++ from the file where it lies (see oracle/ref/gen_shader.py and
 * oracle/Makefile target `ref`): vec2/vec3/vec4/mat3/mat4x4 with the swizzles the shader
 * writes, the built-in functions it calls, samplerCube/sampler1D, gl_FragCoord.
 *
 * Everything here is the PLATFORM — what a GL driver supplies — not the reference's
 * algorithm.  Built-ins map to include/rmdf_glsl.h (`rg_*`, fixed IEEE binary32 operation
 * sequences; conventions C0-C3 of DESIGN.md §2), vector operators are component-wise in
 * the order the GLSL 3.30 specification writes them, texture() goes to the sampler the
 * harness installs (GL fixed function, conventions C4/C5).  The translation unit is built
 * with -ffp-contract=off -fsingle-precision-constant (GLSL literals are 32-bit).
 *
 * This file is included INSIDE a namespace (after <math.h>) so that its sin/cos/pow/...
 * hide libm's overloads for the shader text that follows in the same namespace.
 */
struct vec2; struct vec3; struct vec4;

/* read-only swizzle views living in a union with the components */
template <int N, int A, int B> struct sw2 { float d[N]; inline operator vec2() const; };
template <int N, int A, int B, int C> struct sw3 { float d[N]; inline operator vec3() const; };

struct vec2 {
    float x, y;
    vec2() : x(0.0f), y(0.0f) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float x_, float y_) : x(x_), y(y_) {}
};
struct vec3 {
    union {
        struct { float x, y, z; };
        sw2<3, 0, 1> xy; sw2<3, 0, 2> xz;
        sw3<3, 0, 1, 2> xyz; sw3<3, 2, 0, 1> zxy; sw3<3, 1, 2, 0> yzx;
    };
    vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    vec3(const vec3 &o) : x(o.x), y(o.y), z(o.z) {}
    vec3 &operator=(const vec3 &o) { float a = o.x, b = o.y, c = o.z; x = a; y = b; z = c; return *this; }
};
struct vec4 {
    union {
        struct { float x, y, z, w; };
        sw2<4, 0, 1> xy;
        sw3<4, 0, 1, 2> xyz;
    };
    vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    vec4(vec2 a, float z_, float w_) : x(a.x), y(a.y), z(z_), w(w_) {}
    vec4(vec3 a, float w_) : x(a.x), y(a.y), z(a.z), w(w_) {}
    vec4(const vec4 &o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4 &operator=(const vec4 &o) { float a = o.x, b = o.y, c = o.z, d = o.w; x = a; y = b; z = c; w = d; return *this; }
};
template <int N, int A, int B> inline sw2<N, A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int N, int A, int B, int C> inline sw3<N, A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }

/* ---- operators: component-wise (GLSL 3.30 §5.9) ------------------------------------- */
static inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
static inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
static inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
static inline vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }
static inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
static inline vec2 operator*(float s, vec2 a) { return vec2(s * a.x, s * a.y); }
static inline vec2 operator-(vec2 a, float s) { return vec2(a.x - s, a.y - s); }
static inline vec2 operator+(vec2 a, float s) { return vec2(a.x + s, a.y + s); }

static inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
static inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
static inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
static inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
static inline vec3 operator+(float s, vec3 a) { return vec3(s + a.x, s + a.y, s + a.z); }
static inline vec3 operator-(vec3 a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
static inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
static inline vec3 &operator+=(vec3 &a, vec3 b) { a = a + b; return a; }
static inline vec3 &operator/=(vec3 &a, float s) { a = a / s; return a; }

/* ---- scalar built-ins (include/rmdf_glsl.h) -------------------------------------------- */
static inline float sin(float x) { return rg_sin(x); }
static inline float cos(float x) { return rg_cos(x); }
static inline float tan(float x) { return rg_tan(x); }
static inline float acos(float x) { return rg_acos(x); }
static inline float atan(float y, float x) { return rg_atan2(y, x); }
static inline float exp(float x) { return rg_exp(x); }
static inline float log(float x) { return rg_log(x); }
static inline float sqrt(float x) { return rg_sqrt(x); }
static inline float inversesqrt(float x) { return rg_inversesqrt(x); }
static inline float abs(float x) { return rg_abs(x); }
static inline float floor(float x) { return rg_floor(x); }
static inline float mod(float x, float y) { return rg_mod(x, y); }
static inline float min(float a, float b) { return rg_min(a, b); }
static inline float max(float a, float b) { return rg_max(a, b); }
static inline float clamp(float x, float lo, float hi) { return rg_clamp(x, lo, hi); }
static inline float radians(float d) { return rg_radians(d); }
static inline float mix(float a, float b, float t) { return rg_mix(a, b, t); }
/* pow: convention C1 — a GLSL compiler sees `power = 8` as a constant and may expand
 * pow(r, 7.0) (fragment.shd:148, POWER8) into multiplications; we pick ((r²·r²)·r²)·r.
 * Every other exponent: exp(y·log x), x ≤ 0 → 0 (undefined in GLSL 3.30 §8.2). */
static inline float pow(float x, float y)
{
    if (y == 7.0f) { float x2 = x * x, x4 = x2 * x2; return (x4 * x2) * x; }
    return rg_pow(x, y);
}

/* ---- geometric / vector built-ins (GLSL 3.30 §8.4; conventions C2) -------------------- */
static inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
static inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float length(vec2 a) { return rg_sqrt(dot(a, a)); }
static inline float length(vec3 a) { return rg_sqrt(dot(a, a)); }
static inline float distance(vec3 a, vec3 b) { return length(a - b); }
static inline vec3 normalize(vec3 a) { return a * rg_inversesqrt(dot(a, a)); }
static inline vec2 normalize(vec2 a) { return a * rg_inversesqrt(dot(a, a)); }
static inline vec3 cross(vec3 a, vec3 b)
{
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline vec3 reflect(vec3 i, vec3 n) { float d = 2.0f * dot(n, i); return vec3(i.x - d * n.x, i.y - d * n.y, i.z - d * n.z); }
static inline vec3 abs(vec3 a) { return vec3(rg_abs(a.x), rg_abs(a.y), rg_abs(a.z)); }
static inline vec3 max(vec3 a, float b) { return vec3(rg_max(a.x, b), rg_max(a.y, b), rg_max(a.z, b)); }
static inline vec3 min(vec3 a, float b) { return vec3(rg_min(a.x, b), rg_min(a.y, b), rg_min(a.z, b)); }
static inline vec3 pow(vec3 a, vec3 b) { return vec3(pow(a.x, b.x), pow(a.y, b.y), pow(a.z, b.z)); }
static inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
/* screen-space derivatives: only normal_screen_space_isec (fragment.shd:491-494) uses them and
 * nothing calls it; they exist so that the text compiles */
static inline vec3 dFdx(vec3) { return vec3(0.0f); }
static inline vec3 dFdy(vec3) { return vec3(0.0f); }

/* ---- matrices: column-major constructors (GLSL 3.30 §5.4.2) ----------------------------- */
struct mat4x4 {
    vec4 c[4];
    mat4x4() {}
    mat4x4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3,
           float c0, float c1, float c2, float c3, float d0, float d1, float d2, float d3)
    { c[0] = vec4(a0, a1, a2, a3); c[1] = vec4(b0, b1, b2, b3); c[2] = vec4(c0, c1, c2, c3); c[3] = vec4(d0, d1, d2, d3); }
};
struct mat3 {
    vec3 c[3];
    explicit mat3(const mat4x4 &m) { for (int i = 0; i < 3; i++) c[i] = vec3(m.c[i].x, m.c[i].y, m.c[i].z); }
};
/* matrix · vector: columns scaled and summed left to right (convention C3) */
static inline vec3 operator*(const mat3 &m, vec3 v)
{
    return vec3((m.c[0].x * v.x + m.c[1].x * v.y) + m.c[2].x * v.z,
                (m.c[0].y * v.x + m.c[1].y * v.y) + m.c[2].y * v.z,
                (m.c[0].z * v.x + m.c[1].z * v.y) + m.c[2].z * v.z);
}
static inline vec4 operator*(const mat4x4 &m, vec4 v)
{
    return vec4(((m.c[0].x * v.x + m.c[1].x * v.y) + m.c[2].x * v.z) + m.c[3].x * v.w,
                ((m.c[0].y * v.x + m.c[1].y * v.y) + m.c[2].y * v.z) + m.c[3].y * v.w,
                ((m.c[0].z * v.x + m.c[1].z * v.y) + m.c[2].z * v.z) + m.c[3].z * v.w,
                ((m.c[0].w * v.x + m.c[1].w * v.y) + m.c[2].w * v.z) + m.c[3].w * v.w);
}

/* ---- samplers and fragment inputs: installed by the harness (ref_driver.cpp) -------- */
struct samplerCube { int slot; };
struct sampler1D { const float *texels; };               /* RGB32F, texelFetch = array read */
static thread_local vec4 gl_FragCoord;
static vec4 ref_texture_cube(int slot, vec3 dir);          /* defined by the harness */
static inline vec4 texture(samplerCube s, vec3 dir) { return ref_texture_cube(s.slot, dir); }
static inline vec4 texelFetch(sampler1D s, int i, int /*lod*/)
{
    return vec4(s.texels[i * 3], s.texels[i * 3 + 1], s.texels[i * 3 + 2], 1.0f);
}
