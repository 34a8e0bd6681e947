// frame_sync.cuh — the work queue and the completion signal shared by the march and Julia kernels.
//
// One GPU: the queue counter lives in the context's own memory and nothing is signalled.
// Several GPUs, one process each (parallel.py SharedFrame): a control block on rank 0 — a work-queue word, a "go"
// word and one arrival flag per rank — that every rank maps through CUDA IPC.  All ranks launch the SAME frame-wide
// work list and store their resolved pixels straight into rank 0's frame buffer; the last CTA of each rank's kernel
// publishes the frame's sequence number in the rank's arrival flag and rank 0's stream waits on the flags
// (wait_flags_kernel): no NCCL call and no host round trip on the data path.  Who renders which item:
//   interleaved (default)  groups of a few x-adjacent items are dealt round-robin (group g belongs to rank g mod world);
//                          every rank walks its own items through its LOCAL queue.  Items are 32×8-sample blocks,
//                          neighbouring blocks cost nearly the same, so the comb of every world-th group is an even
//                          share (measured, N = 8: see DESIGN.md §6) — and there is not one remote atomic.
//   shared queue           all ranks claim from the ONE queue word with system-scope atomics over NVLink (a dynamic
//                          deal).  Measured: one address takes ~0.8 G system atomics/s from 8 GPUs — the Julia frame's
//                          2 M claims held N = 4 and N = 8 at the same 2.7 ms, the Mandelbulb frame lost 7 % against the
//                          static deal at N = 8 — and claiming several items at once is no way out, because a warp that
//                          walks k neighbouring items is as slow as one k-times larger item (+8 % for k = 8 on one GPU,
//                          profiles/r2_item_size.md).  Kept as an option.
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
    unsigned int  item_stride;         // interleaved deal, in groups of item_group x-adjacent items: local item j = (q, m) with
    unsigned int  item_offset;         // q = j / item_group, m = j mod item_group is frame item (q·item_stride + item_offset)·
    unsigned int  item_group;          // item_group + m.  A group covers >= 128 B of every output row, so that no 64 B line of
    unsigned int  item_rotate;         // 1: the rank's slot inside each window of item_stride groups rotates with the window (frame_item)
    unsigned int  items_total;         // a HOST frame buffer is shared by two GPUs (partial-line writes over PCIe: the end-to-
                                       // end frame took 4.2 instead of 2.7 ms at N = 8 with single-item interleaving)
    // cost-ordered queue (see "work order" below); both nullptr = items in their natural order, nothing recorded
    const unsigned int *order;         // queue position -> local item, most expensive first (from the previous frame's costs)
    unsigned int *rec;                 // per local item of THIS frame: SM cycles it took
    unsigned int *counts;              // [chunk of kOrderChunk items][cost class] of this frame (zeroed with the queue)
    unsigned int  first_wave;          // queue positions handed out without a claim: position w to warp w of the grid (0 = none)
};

// ---- work order ---------------------------------------------------------------------------------------------------
// A frame ends one worst-case item after the queue runs dry: with items in scan order the warps that claimed the last
// crevice items are still marching (~0.25 ms for a 32×8 item of the Mandelbulb) while everybody else idles — 1.5 % of
// a 15 ms frame, 10 % of one GPU's eighth of it.  A renderer draws frames in sequence and the camera moves slowly, so the
// kernel times every item (SM clock; one store and one fire-and-forget atomic by lane 0: the item's cost, and the
// count of its cost class in its chunk of the item list); one small kernel after the frame (order_place_kernel)
// finishes the counting sort — classes = octaves of the cost, scan order inside a class — and the NEXT frame with the
// same work list walks the queue through that permutation: expensive items first, the cheapest (sky) last, so the
// frame's end is one CHEAP item long.  Pure scheduling: which warp renders an item never changes what it stores.
constexpr int kOrderBuckets = 32;                      // one class per octave of cycles
constexpr int kOrderChunk   = 2048;                    // items per row of the count matrix = per CTA of order_place_kernel
constexpr unsigned kOrderSyncHeader = 32;              // words before the count matrix in an ordered frame's control buffer (one 128 B line)

__device__ __forceinline__ unsigned item_clock()
{
    unsigned t;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(t));
    return t;
}
// queue position -> local item
__device__ __forceinline__ unsigned ordered_item(const FrameSync &S, unsigned pos)
{
    return S.order ? __ldg(S.order + pos) : pos;
}
// Lane 0 notes which item the warp has claimed and when — in two words of shared memory, not in registers: kept live across
// the whole item they cost the march kernel's 64-register budget one instruction per Mandelbulb iteration (33 instead of 32
// issue slots, +3 % on the frame) — and, when the item is finished, what it cost (class 0 = the most expensive octave).
__device__ __forceinline__ void begin_item(const FrameSync &S, unsigned *slot, unsigned item, int lane)
{
    if (S.rec != nullptr && lane == 0) { slot[0] = item; slot[1] = item_clock(); }
}
__device__ __forceinline__ void record_item(const FrameSync &S, const unsigned *slot, int lane)
{
    if (S.rec != nullptr && lane == 0) {
        const unsigned item = slot[0], cycles = item_clock() - slot[1];
        S.rec[item] = cycles;
        atomicAdd(S.counts + (item / kOrderChunk) * kOrderBuckets + __clz((int)(cycles | 1u)), 1u);
    }
}

// local item → frame item of the interleaved deal (identity for a single GPU); >= items_total: nothing to do
__device__ __forceinline__ unsigned frame_item(const FrameSync &S, unsigned j)
{
    if (S.item_stride == 1u) return j;
    const unsigned q = j / S.item_group, m = j - q * S.item_group;
    // window q = the groups q·world … q·world + world − 1, one for every rank: this rank's slot in it rotates with a hash
    // of q (item_rotate).  With a fixed slot a rank owns the same columns of every item row — vertical stripes, and the
    // ranks' kernels differed by 6 % at N = 8 (2.04 … 2.16 ms) — rotated, every rank samples every column
    const unsigned rot = S.item_rotate ? ((q * 2654435761u) >> 16) % S.item_stride : 0u;
    unsigned slot = S.item_offset + rot;
    if (slot >= S.item_stride) slot -= S.item_stride;
    return (q * S.item_stride + slot) * S.item_group + m;
}

// The warp's first queue position, and its next ones (values >= the item count: the queue is drained).  The item loop is
//     for (unsigned q = first_item(S, lane); q < n_items; q = claim_item(S, lane)) …
// with ONE claim site and no state besides q: a stateful "next item" with a first-call test cost the Cornell kernel
// eight registers (72 → 80: 7 → 6 resident CTAs, +4 %).
// A warp's first item needs no claim: grid warp w takes queue position w, the queue word hands out first_wave + 0, 1, …
// With a cost-ordered queue this also PLACES the frame's most expensive items: four per CTA, one per scheduler, CTA
// after CTA over the SMs.  Claimed in arrival order they pile up on some SMs, where eight crevice warps then share one
// scheduler (config 1: 0.111 ms claimed, 0.081 ms placed).
__device__ __forceinline__ unsigned claim_item(const FrameSync &S, int lane)
{
    unsigned base = 0u;
    if (lane == 0) base = S.queue_sys ? atomicAdd_system(S.queue, 1u) : atomicAdd(S.queue, 1u);
    return __shfl_sync(0xffffffffu, base, 0) + S.first_wave;
}
// PLACED = the kernel takes part in the work order (the Mandelbulb scenes).  The others are compiled without the first
// wave and without the cost record: a second source for the loop's queue position alone took the Cornell kernel from 72
// to 80 registers (7 → 6 resident CTAs, +4 %), and they have nothing to gain from an order.
template <bool PLACED>
__device__ __forceinline__ unsigned first_item(const FrameSync &S, int lane)
{
    if (PLACED && S.first_wave != 0u) return blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    return claim_item(S, lane);
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
