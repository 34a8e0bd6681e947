// frame_sync.cu — the two one-CTA kernels that order a multi-GPU frame on the streams (frame_sync.cuh):
//   frame_begin_kernel  rank 0: zero the shared work queue, then publish the frame's sequence number ("go")
//   wait_flags_kernel   any rank: spin until every watched word has reached a sequence number — rank 0 on the
//                       arrival flags of all ranks (the frame is complete), the others on "go" (the queue is reset)
// Spins are bounded: after timeout_ms the kernel gives up and records which word it was waiting for, so a rank that
// died cannot hang the others' GPUs.
#include <cuda_runtime.h>
#include <stdint.h>
#include "rmdf_internal.h"

namespace rmdf {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned int *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void frame_begin_kernel(unsigned int *ctrl, unsigned int seq)
{
    if (threadIdx.x == 0) {
        ctrl[kSyncQueueWord] = 0u;
        __threadfence_system();
        atomicExch_system(ctrl + kSyncGoWord, seq);
    }
}

// sequence numbers only grow: "reached" = (int)(word − value) >= 0, which also survives wrap-around
__global__ void wait_flags_kernel(const unsigned int *flags, int n, unsigned int value, unsigned int timeout_ms, unsigned int *status)
{
    const int i = threadIdx.x;
    if (i >= n) return;
    const unsigned long long t0 = global_timer_ns(), limit = (unsigned long long)timeout_ms * 1000000ull;
    while ((int)(ld_acquire_sys(flags + i) - value) < 0) {
        if (global_timer_ns() - t0 > limit) { atomicExch(status, (unsigned)(i + 1)); return; }
        __nanosleep(100);
    }
}

// CUDA loads kernels lazily, and loading one can need the device to drain — which a spinning wait kernel on the same
// device (several ranks on one GPU: tests, or a host that drives two contexts) would block until its time-out.
// rmdf_create loads both kernels up front.
cudaError_t preload_frame_sync_kernels()
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, frame_begin_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, wait_flags_kernel);
    return e;
}

cudaError_t launch_frame_begin(unsigned int *ctrl, unsigned int seq, cudaStream_t s)
{
    frame_begin_kernel<<<1, 32, 0, s>>>(ctrl, seq);
    return cudaGetLastError();
}
cudaError_t launch_wait_flags(const unsigned int *flags, int n, unsigned int value, unsigned int timeout_ms, unsigned int *status, cudaStream_t s)
{
    wait_flags_kernel<<<1, 32, 0, s>>>(flags, n, value, timeout_ms, status);
    return cudaGetLastError();
}

}  // namespace rmdf
