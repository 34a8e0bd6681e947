"""ray-marching-distance-fields_b200 — B200 (sm_100a) drop-in for the sphere-tracing hot
path of blitzcode/ray-marching-distance-fields.  See DESIGN.md.

Layout:  csrc/ CUDA kernels + C ABI (include/rmdf.h) · host.py host-side mirror of the
reference interface · parallel.py tile sharding over torch.distributed/NCCL ·
haskell/ the `foreign import ccall` shim for the reference's own host code.
"""
from . import _capi
from ._capi import (Params, Counters, RmdfError, SCENE_CORNELL_BOX, SCENE_DE_TEST, SCENE_MB_POWER8,
                    SCENE_MB_GENERAL, ENV_REFLECTION, ENV_COS_1, ENV_COS_8, ENV_COS_64, ENV_COS_512,
                    MODE_PARITY, MODE_FAST, OUT_RGBA8, OUT_RGBA32F, LOD_QUAD, LOD_ALWAYS_LINEAR,
                    LAYOUT_FRAME, LAYOUT_PACKED_TILES, FLAG_NORMAL_CENTRAL, FLAG_CAM_ORTHO, FLAG_RAY_TRACING_TEST, FLAG_SOFT_SHADOWS,
                    DEAL_INTERLEAVED, DEAL_QUEUE)
from .host import (FragmentShader, FrameBuffer, ShaderRenderer, withShaderRenderer, loadHDRImage,
                   saveRadianceImage, decodeHDR, encodeHDR, isTileIdxFirstTile, isTileIdxLastTile,
                   tilesX, tilesY, nTiles)

__all__ = [n for n in dir() if not n.startswith("_")]
