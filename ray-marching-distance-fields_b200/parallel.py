"""Multi-GPU frame rendering: one process per GPU, tiles of the reference's 8×8 grid
(ShaderRendering.hs:49-52) dealt to ranks by (tx + ty) mod G, each rank renders its tiles
packed back to back, and rank 0 gathers them over NCCL (torch.distributed) and reassembles
the frame.  There is no data-path collective besides that gather."""
from __future__ import annotations

import numpy as np


def tiles_of_rank(rank: int, count: int):
    """Tile ids (ty·8 + tx) rendered by `rank`, ascending — the packed order."""
    return [t for t in range(64) if ((t & 7) + (t >> 3)) % count == rank]


def tile_permutation(count: int) -> np.ndarray:
    """perm[t] = index of tile t inside the concatenation of all ranks' packed outputs."""
    perm = np.empty(64, np.int64)
    pos = 0
    for r in range(count):
        for t in tiles_of_rank(r, count):
            perm[t] = pos
            pos += 1
    return perm


def assemble_tiles_numpy(packed_per_rank, w: int, h: int, count: int) -> np.ndarray:
    """packed_per_rank[r]: flat uint32 array of rank r's tiles → [h, w] frame (row 0 = bottom)."""
    tw, th = w // 8, h // 8
    allt = np.concatenate([np.asarray(p).reshape(-1, th, tw) for p in packed_per_rank], axis=0)
    tiles = allt[tile_permutation(count)]                      # [64, th, tw] in tile-id order
    return tiles.reshape(8, 8, th, tw).transpose(0, 2, 1, 3).reshape(h, w)


def assemble_tiles_torch(gathered, w: int, h: int, count: int):
    """Same on the GPU: gathered = [count, tiles_per_rank·th·tw] int32 tensor on rank 0."""
    import torch
    tw, th = w // 8, h // 8
    perm = torch.as_tensor(tile_permutation(count), device=gathered.device)
    tiles = gathered.reshape(64, th, tw).index_select(0, perm)
    return tiles.reshape(8, 8, th, tw).permute(0, 2, 1, 3).reshape(h, w)


def gather_frame(packed, w: int, h: int, world: int, rank: int, gather_list=None):
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
    return assemble_tiles_torch(torch.stack(gather_list), w, h, world)
