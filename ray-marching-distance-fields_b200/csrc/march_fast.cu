// march_fast.cu — fast build of the march kernel.
// Compiled with --use_fast_math (FMA contraction, MUFU sqrt/rsqrt/rcp/lg2/ex2/sin/cos,
// flush-to-zero) and hardware seamless cube-map filtering.  Not bit-identical to the
// oracle: its distance from the parity build is measured against the intrinsic noise
// floor of the eps = 1e-5 normal in tests/test_march_parity.py.
#include "march_kernel.cuh"

namespace rmdf {

struct MathFast {
    static constexpr bool kFast = true;
#ifndef RMDF_FAST_MIN_CTAS
#define RMDF_FAST_MIN_CTAS 6      // x 4 warps = 24 warps/SM, 80 registers
#endif
#ifndef RMDF_FAST_DE_NOINLINE
#define RMDF_FAST_DE_NOINLINE 0
#endif
#ifndef RMDF_FAST_MARCH_NOINLINE
#define RMDF_FAST_MARCH_NOINLINE 0
#endif
    static constexpr bool kNoinlineMarch = RMDF_FAST_MARCH_NOINLINE;
#ifndef RMDF_FAST_MB8_MIN_CTAS
#define RMDF_FAST_MB8_MIN_CTAS 7  // power-8 Mandelbulb: 64 registers, 28 warps/SM (17.7 -> 17.3 ms on config 4)
#endif
    static constexpr int  min_ctas(int scene, bool opt) { return (scene == RMDF_SCENE_MB_POWER8 && !opt) ? RMDF_FAST_MB8_MIN_CTAS : RMDF_FAST_MIN_CTAS; }
    static constexpr bool kNoinlineDE = RMDF_FAST_DE_NOINLINE;
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float inversesqrt(float x) { return rsqrtf(x); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
    static __device__ __forceinline__ float log(float x) { return __logf(x); }
    static __device__ __forceinline__ float exp(float x) { return __expf(x); }
    static __device__ __forceinline__ float pow(float x, float y) { return (x > 0.0f) ? __powf(x, y) : ((x != x) ? x : 0.0f); }
    static __device__ __forceinline__ float sin(float x) { return sinf(x); }
    static __device__ __forceinline__ float cos(float x) { return cosf(x); }
    static __device__ __forceinline__ float acos(float x) { return acosf(x); }
    static __device__ __forceinline__ float atan2(float y, float x) { return atan2f(y, x); }
};

cudaError_t launch_march_fast(const MarchParams &p, int grid, cudaStream_t s) { return launch_march_impl<MathFast>(p, grid, s); }
int march_fast_ctas_per_sm(int scene, int flags) { return march_ctas_per_sm_impl<MathFast>(scene, flags); }

}  // namespace rmdf

#if RMDF_DEBUG_TAIL
extern "C" int rmdf_debug_tail(unsigned long long *out, int n)
{
    return (int)cudaMemcpyFromSymbol(out, rmdf::g_dbg_tail, sizeof(unsigned long long) * (size_t)n);
}
#endif
