// frame_sync.cu — the small kernels around a frame (frame_sync.cuh): order_scatter_kernel (work order for the next
// frame) and the two one-CTA kernels that order a multi-GPU frame on the streams:
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

// ---- work order (frame_sync.cuh): a stable counting sort of the items by cost class, finished in one small launch ------
// The frame's kernel has already counted: counts[chunk][class] for chunks of kOrderChunk = 2048 consecutive items (one
// RED per finished item).  One CTA per chunk: it sums the count matrix (128 B per chunk, from L2) into the class totals
// and the counts of all chunks before its own, its eight warps count their 256 items each, and every item of class b
// goes to start(b) + before(chunk, b) + before(warp, b) + its rank among the warp's items of that class — scan order
// inside a class.  Ranks inside a group of 32 items come from __match_any_sync.
constexpr int kOrderCtaThreads = 256, kOrderCtaWarps = kOrderCtaThreads / 32, kOrderGroups = kOrderChunk / kOrderCtaThreads;

// `sync` is the ordered frame's control buffer: [0] queue word, [1] CTA-exit counter of the frame's kernel, [2] CTA
// counter of this kernel, from word kOrderSyncHeader on the count matrix.  The last CTA to have read the matrix zeroes all
// of it: the next ordered frame finds its queue and counters cleared without a memset of its own.
__global__ void __launch_bounds__(kOrderCtaThreads)
order_place_kernel(const unsigned int *rec, unsigned int n_items, unsigned int *sync, unsigned int n_chunks, unsigned int *order)
{
    __shared__ unsigned int s_before[kOrderCtaWarps][kOrderBuckets], s_total[kOrderCtaWarps][kOrderBuckets];
    __shared__ unsigned int s_cnt[kOrderCtaWarps][kOrderBuckets], s_run[kOrderCtaWarps][kOrderBuckets];
    __shared__ unsigned int s_last;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned lo = blockIdx.x * (unsigned)kOrderChunk + w * (unsigned)(kOrderChunk / kOrderCtaWarps);
    const unsigned int *counts = sync + kOrderSyncHeader;
    // launched with programmatic stream serialization (launch latency overlaps the frame kernel's end): everything the
    // frame's kernel wrote is visible after this
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // this warp's items: all loads first (the loops below are chains of warp votes)
    int cls[kOrderGroups];
#pragma unroll
    for (int g = 0; g < kOrderGroups; g++) {
        const unsigned i = lo + 32u * g + lane;
        cls[g] = i < n_items ? __clz((int)(__ldcg(rec + i) | 1u)) : 32;      // class 0 = the most expensive octave; 32 = no item
    }
    // thread (w, lane) sums class `lane` over the chunks w, w + 8, …: eight independent loads in flight per thread
    {
        unsigned before = 0u, total = 0u;
        for (unsigned r0 = w; r0 < n_chunks; r0 += 8u * kOrderCtaWarps) {
            unsigned c[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const unsigned r = r0 + (unsigned)u * kOrderCtaWarps;
                c[u] = r < n_chunks ? __ldcg(counts + r * kOrderBuckets + lane) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                total += c[u];
                if (r0 + (unsigned)u * kOrderCtaWarps < blockIdx.x) before += c[u];
            }
        }
        s_before[w][lane] = before; s_total[w][lane] = total;
    }
    s_cnt[w][lane] = 0u;
    __syncwarp();
#pragma unroll
    for (int g = 0; g < kOrderGroups; g++) {
        const unsigned same = __match_any_sync(0xffffffffu, cls[g]);
        if (cls[g] < 32 && (same & lt) == 0u) s_cnt[w][cls[g]] += (unsigned)__popc(same);   // one leader per class, distinct words
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                                 // this CTA is through with the matrix
        __threadfence();
        s_last = (atomicAdd(sync + 2, 1u) == gridDim.x - 1u) ? 1u : 0u;
    }
    {
        unsigned before = 0u, total = 0u;
        for (int j = 0; j < kOrderCtaWarps; j++) { before += s_before[j][lane]; total += s_total[j][lane]; }
        for (unsigned j = 0; j < w; j++) before += s_cnt[j][lane];
        unsigned incl = total;                                              // exclusive scan of the class totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += v;
        }
        s_run[w][lane] = incl - total + before;                             // first position of this warp's items of class `lane`
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < kOrderGroups; g++) {
        const unsigned same = __match_any_sync(0xffffffffu, cls[g]);
        if (cls[g] < 32) {
            const unsigned pos = s_run[w][cls[g]] + (unsigned)__popc(same & lt);
            if (pos < n_items) order[pos] = lo + 32u * g + lane;            // always true for a completed frame; never write outside
        }
        __syncwarp();
        if (cls[g] < 32 && (same & lt) == 0u) s_run[w][cls[g]] += (unsigned)__popc(same);
        __syncwarp();
    }
    __syncthreads();
    if (s_last) {                                                           // every CTA has read the matrix: clear it, and the header
        uint4 *p = reinterpret_cast<uint4 *>(sync);
        const unsigned n16 = (kOrderSyncHeader + n_chunks * kOrderBuckets) / 4u;
        for (unsigned i = threadIdx.x; i < n16; i += kOrderCtaThreads) p[i] = make_uint4(0u, 0u, 0u, 0u);
    }
}
cudaError_t launch_order_place(const unsigned int *rec, unsigned int n_items, unsigned int *sync, unsigned int *order, cudaStream_t s)
{
    const unsigned n_chunks = (n_items + kOrderChunk - 1u) / kOrderChunk;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_chunks); cfg.blockDim = dim3(kOrderCtaThreads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, order_place_kernel, rec, n_items, sync, n_chunks, order);
}

// CUDA loads kernels lazily, and loading one can need the device to drain — which a spinning wait kernel on the same
// device (several ranks on one GPU: tests, or a host that drives two contexts) would block until its time-out.
// rmdf_create loads both kernels up front.
cudaError_t preload_frame_sync_kernels()
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, frame_begin_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, wait_flags_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, order_place_kernel);
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
