/* rmdf_glsl.h — deterministic FP32 definition of the GLSL 3.30 built-ins that
 * the reference's fragment shader (/root/reference/fragment.shd) calls.
 *
 * WHY THIS EXISTS.  GLSL leaves the precision of pow/log/exp/sin/cos/acos/atan/
 * inversesqrt/normalize to the implementation, and the reference's sphere tracer
 * is chaotic in those roundings: `t += dist` feeds the next march position and the
 * surface normal is a finite difference with eps = 1e-5 (fragment.shd:463-470), so
 * a 1-ulp change in `log(r)` re-randomises ~0.2 % of the normals.
 * This header pins every built-in to a fixed sequence of IEEE-754 binary32
 * operations (+ - * / sqrt, compares, integer bit moves) and nothing else, so that
 * a CPU compiled with `-ffp-contract=off` and a GPU compiled with `-fmad=false`
 * produce the SAME BITS.  It is the "platform" (the stand-in for the GL driver's
 * built-in library, which is not part of the reference repository), shared by
 *   - the CUDA march kernel's parity build (csrc/march_parity.cu), and
 *   - the CPU oracle's restatement of fragment.shd (oracle/oracle_march.c).
 * The algorithm itself (distance estimators, march loop, normals, AO, shading) is
 * written twice, independently, on top of it.
 *
 * Accuracy (checked by tests/test_glsl_builtins.py against libm in double):
 *   log, exp ≤ 1 ulp; sin, cos ≤ 2 ulp for |x| < 8192; atan2, acos ≤ 3 ulp.
 * Valid C99 and CUDA C++.  No FMA is ever written explicitly; the including
 * translation unit must disable contraction for bit-parity.
 */
#ifndef RMDF_GLSL_H
#define RMDF_GLSL_H

#include <stdint.h>

#if defined(__CUDACC__)
#  define RG_HD static __host__ __device__ __forceinline__
#  include <math.h>
#  include <string.h>
#else
#  include <math.h>
#  include <string.h>
#  define RG_HD static inline
#endif

/* ---- bit moves ---------------------------------------------------------- */
RG_HD int32_t rg_f2i(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int32_t i; memcpy(&i, &f, 4); return i;
#endif
}
RG_HD float rg_i2f(int32_t i)
{
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}

/* ---- scalar built-ins --------------------------------------------------- */
/* Comparison-based so that NaN behaviour is identical on every target:
 * min(a,b) = (b < a) ? b : a ; max(a,b) = (a < b) ? b : a   (GLSL 3.30 §8.3) */
RG_HD float rg_min(float a, float b) { return (b < a) ? b : a; }
RG_HD float rg_max(float a, float b) { return (a < b) ? b : a; }
RG_HD float rg_clamp(float x, float lo, float hi) { return rg_min(rg_max(x, lo), hi); }
RG_HD float rg_abs(float x) { return rg_i2f(rg_f2i(x) & 0x7fffffff); }
RG_HD float rg_floor(float x) { return floorf(x); }
RG_HD float rg_mod(float x, float y) { return x - y * rg_floor(x / y); }
RG_HD float rg_sqrt(float x) { return sqrtf(x); }                 /* IEEE sqrt.rn  */
RG_HD float rg_inversesqrt(float x) { return 1.0f / sqrtf(x); }   /* sqrt.rn, div.rn */
RG_HD float rg_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
RG_HD float rg_radians(float deg) { return deg * 0.017453292519943295f; }

/* log(x), natural.  x = m·2^k, m ∈ [√½, √2); f = m−1; s = f/(2+f);
 * log(1+f) = f − f²/2 + s·(f²/2 + R(s²)).  x ≤ 0 → −inf (x = 0) or NaN. */
RG_HD float rg_log(float x)
{
    int32_t ix = rg_f2i(x);
    int32_t k = 0;
    if (ix < 0x00800000) {                       /* zero, denormal or negative */
        if ((ix & 0x7fffffff) == 0) return rg_i2f((int32_t)0xff800000);  /* -inf */
        if (ix < 0) return rg_i2f(0x7fc00000);                          /* NaN  */
        x = x * 33554432.0f;                     /* 2^25: normalise denormals  */
        ix = rg_f2i(x);
        k = -25;
    }
    if (ix >= 0x7f800000) return x;              /* inf or NaN */
    k += (ix >> 23) - 127;
    ix &= 0x007fffff;
    {
        /* pick m in [sqrt(1/2), sqrt(2)) */
        int32_t i = (ix + 0x4afb0d) & 0x00800000;   /* 1 iff mantissa ≥ √2 */
        float m = rg_i2f(ix | (i ^ 0x3f800000));
        float f, s, z, R, hfsq, dk;
        k += (i >> 23);
        f = m - 1.0f;
        s = f / (2.0f + f);
        z = s * s;
        R = z * (0.6666666865348816f + z * (0.40000131726264954f +
            z * (0.28549882769584656f + z * 0.23355302214622498f)));
        hfsq = 0.5f * f * f;
        dk = (float)k;
        /* ln2 split: hi has 16 bits so dk*hi is exact */
        return dk * 0.693145751953125f -
               ((hfsq - (s * (hfsq + R) + dk * 1.428606765330187e-06f)) - f);
    }
}

/* exp(x).  k = round(x/ln2); r = x − k·ln2; exp(r) = 1 + r + r·c/(2−c),
 * c = r − r²·P(r²).  Overflow → +inf, underflow → denormals → 0. */
RG_HD float rg_exp(float x)
{
    float fk, hi, lo, r, t, c, y;
    int32_t k;
    if (x != x) return x;
    if (x > 88.72283935546875f) return rg_i2f(0x7f800000);
    if (x < -104.0f) return 0.0f;
    fk = rg_floor(x * 1.4426950216293335f + 0.5f);
    k = (int32_t)fk;
    hi = x - fk * 0.693145751953125f;
    lo = fk * 1.428606765330187e-06f;
    r = hi - lo;
    t = r * r;
    c = r - t * (0.1666666716337204f + t * (-0.0027777631767094135f +
        t * 6.5825974161271e-05f));
    y = 1.0f - ((lo - (r * c) / (2.0f - c)) - hi);
    if (k >= -125) {
        if (k > 127) { y = y * 2.0f; k -= 1; }          /* k = 128 edge */
        return y * rg_i2f((k + 127) << 23);
    }
    /* result is denormal: scale in two exact steps */
    return (y * rg_i2f((k + 100 + 127) << 23)) * 7.888609052210118e-31f; /* 2^-100 */
}

/* pow(x,y) = exp(y·log x) for x > 0 (GLSL 3.30 §8.2: undefined for x < 0 and for
 * x = 0 with y ≤ 0; we return 0 for every x ≤ 0, NaN stays NaN). */
RG_HD float rg_pow(float x, float y)
{
    if (x != x) return x;
    if (!(x > 0.0f)) return 0.0f;
    return rg_exp(y * rg_log(x));
}

/* sin / cos: Cody–Waite reduction by π/2 in four pieces (three of 11 bits, so
 * n·piece is exact for |n| < 2^13, i.e. |x| < ~1.2e4), then odd/even polynomials
 * on [−π/4, π/4]. */
RG_HD float rg__reduce_pio2(float x, int32_t *quad)
{
    float fn = rg_floor(x * 0.6366197466850281f + 0.5f);
    float r;
    *quad = (int32_t)fn;
    r = x - fn * 1.5703125f;
    r = r - fn * 0.0004837512969970703f;
    r = r - fn * 7.549533620476723e-08f;
    r = r - fn * 2.5633440682570896e-12f;
    return r;
}
RG_HD float rg__ksin(float r)
{
    float z = r * r;
    return r + r * (z * (-0.1666666716337204f + z * (0.008333331905305386f +
           z * (-0.00019840040476992726f + z * 2.724389105424052e-06f))));
}
RG_HD float rg__kcos(float r)
{
    float z = r * r;
    return (1.0f - 0.5f * z) + (z * z) * (0.0416666679084301f +
           z * (-0.00138888880610466f + z * (2.4800561732263304e-05f +
           z * -2.729591130901099e-07f)));
}
RG_HD float rg_sin(float x)
{
    int32_t q; float r = rg__reduce_pio2(x, &q);
    switch (q & 3) {
        case 0:  return  rg__ksin(r);
        case 1:  return  rg__kcos(r);
        case 2:  return -rg__ksin(r);
        default: return -rg__kcos(r);
    }
}
RG_HD float rg_cos(float x)
{
    int32_t q; float r = rg__reduce_pio2(x, &q);
    switch (q & 3) {
        case 0:  return  rg__kcos(r);
        case 1:  return -rg__ksin(r);
        case 2:  return -rg__kcos(r);
        default: return  rg__ksin(r);
    }
}
RG_HD float rg_tan(float x) { return rg_sin(x) / rg_cos(x); }

/* atan on [0,1] → atan2 → acos. */
RG_HD float rg__atan01(float t)           /* 0 ≤ t ≤ 1 */
{
    float base = 0.0f, z;
    if (t > 0.41421356237309503f) {       /* tan(π/8) */
        t = (t - 1.0f) / (t + 1.0f);
        base = 0.7853981852531433f;       /* π/4 */
    }
    z = t * t;
    return base + (t + t * (z * (-0.3333333134651184f + z * (0.19999542832374573f +
           z * (-0.1426393687725067f + z * (0.10742649435997009f +
           z * -0.0644615963101387f))))));
}
RG_HD float rg_atan2(float y, float x)    /* GLSL atan(y, x) */
{
    float ax = rg_abs(x), ay = rg_abs(y), a;
    if (ax == 0.0f && ay == 0.0f) return 0.0f;
    if (ay > ax) a = 1.5707963705062866f - rg__atan01(ax / ay);
    else         a = rg__atan01(ay / ax);
    if (x < 0.0f) a = 3.1415927410125732f - a;
    return (y < 0.0f) ? -a : a;
}
RG_HD float rg_acos(float x)
{
    return rg_atan2(rg_sqrt((1.0f - x) * (1.0f + x)), x);
}

/* ---- vec3 ----------------------------------------------------------------- */
typedef struct { float x, y, z; } rg_vec3;

RG_HD rg_vec3 rg_v3(float x, float y, float z) { rg_vec3 v; v.x = x; v.y = y; v.z = z; return v; }
RG_HD rg_vec3 rg_add(rg_vec3 a, rg_vec3 b) { return rg_v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RG_HD rg_vec3 rg_sub(rg_vec3 a, rg_vec3 b) { return rg_v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RG_HD rg_vec3 rg_mul(rg_vec3 a, rg_vec3 b) { return rg_v3(a.x * b.x, a.y * b.y, a.z * b.z); }
RG_HD rg_vec3 rg_scale(rg_vec3 a, float s) { return rg_v3(a.x * s, a.y * s, a.z * s); }
RG_HD rg_vec3 rg_neg(rg_vec3 a) { return rg_v3(-a.x, -a.y, -a.z); }
/* dot: products summed left to right, no fused operations */
RG_HD float rg_dot(rg_vec3 a, rg_vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
RG_HD float rg_dot2(float ax, float ay, float bx, float by) { return ax * bx + ay * by; }
RG_HD float rg_length(rg_vec3 a) { return rg_sqrt(rg_dot(a, a)); }
RG_HD float rg_length2(float x, float y) { return rg_sqrt(x * x + y * y); }
RG_HD float rg_distance(rg_vec3 a, rg_vec3 b) { return rg_length(rg_sub(a, b)); }
/* normalize(v) = v · inversesqrt(dot(v,v)) */
RG_HD rg_vec3 rg_normalize(rg_vec3 a) { return rg_scale(a, rg_inversesqrt(rg_dot(a, a))); }
RG_HD rg_vec3 rg_cross(rg_vec3 a, rg_vec3 b)
{
    return rg_v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
/* reflect(I,N) = I − 2·dot(N,I)·N  (GLSL 3.30 §8.4) */
RG_HD rg_vec3 rg_reflect(rg_vec3 i, rg_vec3 n)
{
    float d = 2.0f * rg_dot(n, i);
    return rg_v3(i.x - d * n.x, i.y - d * n.y, i.z - d * n.z);
}

/* ---- fixed-function pieces (OpenGL 3.3 core) -------- */
/* float → UNORM8: clamp to [0,1], round to nearest; NaN → 0. */
RG_HD uint32_t rg_unorm8(float c)
{
    if (!(c > 0.0f)) return 0u;
    if (c > 1.0f) c = 1.0f;
    return (uint32_t)(c * 255.0f + 0.5f);
}
/* IEEE binary16 round-to-nearest-even, as the GL upload to RGB16F does
 * (HDREnvMap.hs:160-161).  Returns the half's value widened back to float. */
RG_HD uint16_t rg_float_to_half_bits(float f)
{
    uint32_t x = (uint32_t)rg_f2i(f);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u));
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);            /* ≥ 65520 → inf */
    if (ax < 0x33000001u) return (uint16_t)sign;                          /* < 2^-25 → 0  */
    if (ax < 0x38800000u) {                                               /* half denormal */
        uint32_t shift = 126u - (ax >> 23);           /* 14 … 24 */
        uint32_t man = (ax & 0x007fffffu) | 0x00800000u;
        uint32_t h = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1u);
        uint32_t halfway = 1u << (shift - 1u);
        if (rem > halfway || (rem == halfway && (h & 1u))) h += 1u;
        return (uint16_t)(sign | h);
    }
    {
        uint32_t h = ((ax - 0x38000000u) >> 13);
        uint32_t rem = ax & 0x1fffu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h += 1u;
        return (uint16_t)(sign | h);
    }
}
RG_HD float rg_half_bits_to_float(uint16_t hb)
{
    uint32_t sign = ((uint32_t)hb & 0x8000u) << 16;
    uint32_t e = ((uint32_t)hb >> 10) & 0x1fu;
    uint32_t m = (uint32_t)hb & 0x3ffu;
    if (e == 0u) {
        float v = (float)m * 5.9604644775390625e-08f;      /* m · 2^-24, exact */
        return rg_i2f((int32_t)(sign | (uint32_t)rg_f2i(v)));
    }
    if (e == 31u) return rg_i2f((int32_t)(sign | 0x7f800000u | (m << 13)));
    return rg_i2f((int32_t)(sign | ((e + 112u) << 23) | (m << 13)));
}
RG_HD float rg_round_to_half(float f) { return rg_half_bits_to_float(rg_float_to_half_bits(f)); }

#endif /* RMDF_GLSL_H */
