// frame_sync.cuh — the work queue and the completion signal shared by the march and Julia kernels.
//
// One GPU: the queue counter lives in the context's own memory and nothing is signalled.
// Several GPUs, one process each (parallel.py SharedFrame): the queue counter, a "go" word and one arrival flag per
// rank live in a control block on rank 0 that every rank maps through CUDA IPC.  All ranks claim work items of the
// SAME frame-wide work list from that one counter with system-scope atomics over NVLink (a dynamic deal: a GPU that
// draws cheap items simply claims more), store their resolved pixels straight into rank 0's frame buffer, and the
// last CTA of each rank's kernel publishes the frame's sequence number in the rank's arrival flag.  Rank 0's stream
// then waits on the flags (wait_flags_kernel): no NCCL call and no host round trip on the data path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rmdf {

constexpr int kSyncQueueWord = 0;      // layout of the control block (uint32 words), see include/rmdf.h RMDF_SYNC_WORDS
constexpr int kSyncGoWord    = 1;
constexpr int kSyncFlagWord0 = 8;      // arrival flag of rank r = word 8 + r
constexpr int kSyncWords     = 64;

struct FrameSync {
    unsigned int *queue;               // work-queue counter (zeroed before the frame)
    unsigned int *done;                // this kernel's CTA-exit counter (context-local, zeroed before the launch)
    unsigned int *signal;              // arrival flag to publish when the last CTA has stored; nullptr = none
    unsigned int  signal_value;
    int           queue_sys;           // the queue is shared with other GPUs: claim with a system-scope atomic
    unsigned int  prefetch_margin;     // 0: claim an item when the previous one is done.  m > 0: while item + m < n_items
                                       // the NEXT item is claimed before working on the current one, which hides the
                                       // NVLink round trip of a shared queue; over the last m items the warps go back to
                                       // claiming late, so the end of the frame is still dealt dynamically (an item
                                       // claimed early is an item an idle warp can no longer take: unconditional
                                       // prefetching cost BASELINE config 1 0.099 → 0.124 ms and 1.4 % at N = 2)
};

// lane 0 claims; the value is broadcast by the caller when it needs it, so the atomic's latency (a NVLink round trip
// for a shared queue) overlaps the work on the previous item
__device__ __forceinline__ unsigned claim_item(const FrameSync &S)
{
    return S.queue_sys ? atomicAdd_system(S.queue, 1u) : atomicAdd(S.queue, 1u);
}

// kernel epilogue, reached by every thread of the CTA: when this was the grid's last CTA, all of the grid's stores
// (also those to peer or host memory) are ordered before the flag becomes visible system-wide
__device__ __forceinline__ void signal_when_grid_done(const FrameSync &S)
{
    if (S.signal == nullptr) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(S.done, 1u);
        if (prev == gridDim.x - 1u) {
            __threadfence_system();
            atomicExch_system(S.signal, S.signal_value);
        }
    }
}

}  // namespace rmdf
