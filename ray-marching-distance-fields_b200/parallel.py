"""Multi-GPU frame rendering: one process per GPU.  The frame is cut into a shard grid of
gx × gy tiles (default: the reference's 8×8 grid, ShaderRendering.hs:49-52); tile (tx, ty)
belongs to rank (tx + skew·ty) mod G, each rank renders its tiles packed back to back in
ascending (ty, tx) order, and rank 0 gathers them with ONE collective (torch.distributed
gather over NCCL) and reassembles the frame.  Samples are independent, so there is no other
data-path communication.

A finer grid with a skew balances the ranks: the Mandelbulb's cost is concentrated mid-frame
(measured with the oracle's counters: 8×8 skew 1 → 90 % at G = 8; 32×16 skew 2 → 97 %; 240×135
skew 4 → 99.5 %, see balanced_grid)."""
from __future__ import annotations

import numpy as np

DEFAULT_GRID = (8, 8)
DEFAULT_SKEW = 1


def balanced_grid(fb_w: int, fb_h: int, resolve_log2: int, world: int, min_tile: int = 64, max_grid: int = 256):
    """Finest shard grid (≤ max_grid per axis, tiles of at least min_tile samples) that keeps tiles on even
    pixel boundaries, a multiple of the resolve footprint and gx a multiple of the world size; skew G/2.
    The finer the grid, the closer the static deal comes to the cost distribution: with the oracle's work
    counters as the cost model, config 4 on 8 ranks balances to 90 % with the reference's 8×8 tiles, 97 % with
    32×16 and 99.5 % with 240×135 (64×64-sample tiles); the kernel walks a rank's tiles as one work list, so
    small tiles cost nothing extra."""
    q = 1 << resolve_log2
    def pick(dim, need_mult):
        for g in range(min(max_grid, dim // max(min_tile, 2)), 0, -1):
            if g % need_mult == 0 and dim % (2 * g) == 0 and (dim // g) % q == 0:
                return g
        raise ValueError("no shard grid fits dimension %d" % dim)
    return (pick(fb_w, world), pick(fb_h, 1)), max(1, world // 2)


def tiles_of_rank(rank: int, count: int, grid=DEFAULT_GRID, skew: int = DEFAULT_SKEW):
    """Tile ids (ty·gx + tx) rendered by `rank`, in packed order."""
    gx, gy = grid
    return [ty * gx + tx for ty in range(gy) for tx in range(gx) if (tx + skew * ty) % count == rank]


def tile_permutation(count: int, grid=DEFAULT_GRID, skew: int = DEFAULT_SKEW) -> np.ndarray:
    """perm[t] = index of tile t inside the concatenation of all ranks' packed outputs."""
    gx, gy = grid
    perm = np.empty(gx * gy, np.int64)
    pos = 0
    for r in range(count):
        for t in tiles_of_rank(r, count, grid, skew):
            perm[t] = pos
            pos += 1
    return perm


def assemble_tiles_numpy(packed_per_rank, w: int, h: int, count: int, grid=DEFAULT_GRID, skew: int = DEFAULT_SKEW) -> np.ndarray:
    """packed_per_rank[r]: flat uint32 array of rank r's tiles → [h, w] frame (row 0 = bottom)."""
    gx, gy = grid
    tw, th = w // gx, h // gy
    allt = np.concatenate([np.asarray(p).reshape(-1, th, tw) for p in packed_per_rank], axis=0)
    tiles = allt[tile_permutation(count, grid, skew)]          # [gy·gx, th, tw] in tile-id order
    return tiles.reshape(gy, gx, th, tw).transpose(0, 2, 1, 3).reshape(h, w)


_perm_cache = {}


def assemble_tiles_torch(gathered, w: int, h: int, count: int, grid=DEFAULT_GRID, skew: int = DEFAULT_SKEW):
    """Same on the GPU: gathered = [count, tiles_per_rank·th·tw] int32 tensor on rank 0."""
    import torch
    gx, gy = grid
    tw, th = w // gx, h // gy
    key = (count, grid, skew, str(gathered.device))
    if key not in _perm_cache:
        _perm_cache[key] = torch.as_tensor(tile_permutation(count, grid, skew), device=gathered.device)
    tiles = gathered.reshape(gx * gy, th, tw).index_select(0, _perm_cache[key])
    return tiles.reshape(gy, gx, th, tw).permute(0, 2, 1, 3).reshape(h, w)


def gather_frame(packed, w: int, h: int, world: int, rank: int, gather_list=None, grid=DEFAULT_GRID, skew: int = DEFAULT_SKEW):
    """Rank-local packed tiles (1-D int32 tensor, CUDA+NCCL or CPU+gloo) → the [h, w] frame on
    rank 0 (None elsewhere).  One gather, no other collective."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return packed.reshape(h, w)
    if rank == 0 and gather_list is None:
        gather_list = [torch.empty_like(packed) for _ in range(world)]
    dist.gather(packed, gather_list if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    return assemble_tiles_torch(torch.stack(gather_list), w, h, world, grid, skew)


# ---- Julia: interleaved row bands (the reference splits rows, Fractal2D.hs:98) -----------------
def julia_band_height(h: int, resolve_log2: int, world: int, min_bands_per_rank: int = 32) -> int:
    """Largest band height that is a multiple of 32 rows (one work-item row) and of the resolve
    footprint, divides h, and leaves every rank at least `min_bands_per_rank` bands."""
    q = max(32, 1 << resolve_log2)
    best = None
    for bh in range(q, h + 1, q):
        if h % bh == 0 and (h // bh) % world == 0 and (h // bh) // world >= min_bands_per_rank:
            best = bh
    if best is None:
        for bh in range(1 << resolve_log2, h + 1, 1 << resolve_log2):
            if h % bh == 0 and (h // bh) % world == 0:
                best = bh if best is None else best
    if best is None:
        raise ValueError("no band height fits h=%d for %d ranks" % (h, world))
    return best


def assemble_bands_numpy(packed_per_rank, w: int, h: int, count: int, band_h: int) -> np.ndarray:
    """packed_per_rank[r]: rank r's bands (b ≡ r mod count, ascending) → [h, w] (resolved sizes)."""
    nb = h // band_h // count
    allb = np.stack([np.asarray(p).reshape(nb, band_h, w) for p in packed_per_rank], axis=1)   # [nb, count, band_h, w]
    return allb.reshape(h, w)


def gather_bands(packed, w: int, h: int, world: int, rank: int, band_h: int, gather_list=None):
    """torch version with one gather to rank 0 (None elsewhere); w, h, band_h in output pixels."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return packed.reshape(h, w)
    if rank == 0 and gather_list is None:
        gather_list = [torch.empty_like(packed) for _ in range(world)]
    dist.gather(packed, gather_list if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    nb = h // band_h // world
    return torch.stack(gather_list).reshape(world, nb, band_h, w).permute(1, 0, 2, 3).reshape(h, w)


# ---- fused gather: every rank's kernel stores its pixels straight into rank 0's frame buffer ----
SYNC_WORDS = 64          # include/rmdf.h RMDF_SYNC_WORDS


class SharedFrame:
    """The resolved frame buffer lives on rank 0's GPU; the other ranks map it through CUDA IPC and pass `self.ptr`
    as the render target, so the kernels' 128-bit stores cross NVLink while the rest of the frame is still being
    marched.  Behind the frame sits a control block (work queue, "go" word, one arrival flag per rank):

      tiles + barrier  renderDevice(params with shard_rank/shard_count, self.ptr); complete()   — round 1: one NCCL barrier
      flags            begin(); renderDevice(the SAME full-frame params on every rank, self.ptr); end()
                       work item i (a 32×8-sample block) belongs to rank i mod world (or, deal = 1, all GPUs claim items
                       from ONE queue with system-scope atomics over NVLink); each kernel's last CTA publishes the frame
                       number in its rank's arrival flag and rank 0's stream waits on the flags: no NCCL call, no host
                       round trip on the data path.
    """

    def __init__(self, sr, n_words: int, world: int, rank: int):
        import torch.distributed as dist
        self.sr, self.world, self.rank, self.n_words = sr, world, rank, n_words
        self.seq = 0
        frame_bytes = (n_words * 4 + 255) & ~255
        box = [None]
        if rank == 0:
            self.ptr, handle = sr.sharedAlloc(frame_bytes + SYNC_WORDS * 4)      # zero-initialised by the library
            box[0] = handle
        if world > 1:
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                self.ptr = sr.sharedOpen(box[0])
        self.ctrl = self.ptr + frame_bytes

    # -- dynamic deal ---------------------------------------------------------------------------------
    def begin(self, timeout_ms: int = 10000, deal: int = 0) -> int:
        """Every rank, before its render call of the frame.  deal: 0 = interleaved (item i → rank i mod world, local
        queues), 1 = one shared queue claimed over NVLink.  Returns the frame's sequence number."""
        self.seq += 1
        self.sr.frameBegin(self.ctrl, self.rank, self.world, self.seq, timeout_ms, deal)
        return self.seq

    def end(self, timeout_ms: int = 10000) -> None:
        """Rank 0: work queued on its stream after this sees the complete frame.  Other ranks: nothing to do."""
        if self.rank == 0:
            self.sr.frameEnd(self.ctrl, self.world, self.seq, timeout_ms)

    # -- static deal ----------------------------------------------------------------------------------
    def complete(self):
        """Static deal: all ranks' stores are in rank 0's memory once this returns.  The renderer may launch on its
        own non-blocking stream, which a NCCL barrier does not order after: synchronise it first."""
        import torch.distributed as dist
        self.sr.sync()
        if self.world > 1:
            dist.barrier()

    def as_tensor(self):
        """Rank 0: a torch int32 view of the frame (no copy)."""
        import torch
        assert self.rank == 0

        class _Wrap:
            pass
        w = _Wrap()
        w.__cuda_array_interface__ = {"shape": (self.n_words,), "typestr": "<i4", "data": (self.ptr, False), "version": 3}
        return torch.as_tensor(w, device="cuda")

    def close(self):
        """Importers close their mapping first; only then does rank 0 free the exported allocation (freeing exported
        memory that is still mapped elsewhere is undefined behaviour)."""
        import torch.distributed as dist
        self.sr.sync()
        if self.rank != 0 and self.world > 1:
            self.sr.sharedClose(self.ptr)
        if self.world > 1:
            dist.barrier()
        if self.rank == 0:
            self.sr.sharedFree(self.ptr)
        self.ptr = None


class SharedHostFrame:
    """The end-to-end variant of SharedFrame: the resolved frame lives in HOST memory (one POSIX shared-memory
    segment that every rank maps and page-locks for its own GPU, `rmdf_host_register`).  Every rank's kernel
    stores its tiles over its own PCIe link straight into the final host frame, so after the launch there is
    one barrier and nothing else — no gather, and no device→host copy on rank 0 (33 MB per config-4 frame)."""

    def __init__(self, sr, n_words: int, world: int, rank: int):
        import ctypes as C
        from multiprocessing import shared_memory
        import torch.distributed as dist
        self.sr, self.world, self.rank, self.n_words = sr, world, rank, n_words
        box = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=n_words * 4)
            box[0] = self.shm.name
        if world > 1:
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                self.shm = shared_memory.SharedMemory(name=box[0])
                try:                 # only the creator unlinks: keep Python's resource tracker from trying again at exit
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(self.shm._name, "shared_memory")
                except Exception:
                    pass
        self.host_ptr = C.addressof(C.c_char.from_buffer(self.shm.buf))
        try:
            self.ptr = sr.hostRegister(self.host_ptr, n_words * 4)
        except BaseException:                # do not leak the /dev/shm segment: this object is never returned
            self.host_ptr = None
            try:
                self.shm.close()
            except BufferError:
                pass
            if rank == 0:
                self.shm.unlink()
            raise

    def complete(self):
        """Every rank: its kernel has finished (stream synchronised) and so have all the others'."""
        import torch
        import torch.distributed as dist
        cuda = torch.cuda.is_available()             # (the CPU test drives this class with gloo and no device)
        if cuda:
            torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            if cuda:
                torch.cuda.synchronize()

    def as_numpy(self):
        """Rank 0: the frame as a uint32 array over the shared segment (no copy)."""
        return np.frombuffer(self.shm.buf, dtype=np.uint32, count=self.n_words)

    def close(self):
        import torch.distributed as dist
        self.sr.hostUnregister(self.host_ptr)
        if self.world > 1:
            dist.barrier()
        self.host_ptr = None
        try:
            self.shm.close()
        except BufferError:          # a numpy view is still alive; the mapping goes with the process
            pass
        if self.rank == 0:
            self.shm.unlink()
