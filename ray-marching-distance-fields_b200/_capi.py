"""ctypes binding of include/rmdf.h (librmdf_b200.so).  No CPU fallback: importing this
module without the built CUDA library raises, and every compute call fails loudly
when there is no sm_100 GPU."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["RMDF_LIB"]) if os.environ.get("RMDF_LIB") else HERE / "librmdf_b200.so"   # RMDF_LIB: tuning variants

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = range(5)
SCENE_CORNELL_BOX, SCENE_DE_TEST, SCENE_MB_POWER8, SCENE_MB_GENERAL = range(4)
ENV_REFLECTION, ENV_COS_1, ENV_COS_8, ENV_COS_64, ENV_COS_512 = range(5)
MODE_PARITY, MODE_FAST = 0, 1
OUT_RGBA8, OUT_RGBA32F = 0, 1
LOD_QUAD, LOD_ALWAYS_LINEAR = 0, 1
LAYOUT_FRAME, LAYOUT_PACKED_TILES = 0, 1
FLAG_NORMAL_CENTRAL, FLAG_CAM_ORTHO, FLAG_RAY_TRACING_TEST, FLAG_SOFT_SHADOWS = 1, 2, 4, 8
DEAL_INTERLEAVED, DEAL_QUEUE = 0, 1

EXPORTS = [
    "rmdf_create", "rmdf_destroy", "rmdf_last_error", "rmdf_ctx_last_error", "rmdf_abi_version", "rmdf_set_stream", "rmdf_sync",
    "rmdf_envmap_resize_height", "rmdf_envmap_resize", "rmdf_envmap_convolve", "rmdf_envmap_convolve_device", "rmdf_envmap_convolve_rows_device",
    "rmdf_rgbe_roundtrip", "rmdf_envmap_bind", "rmdf_envmap_bind_cube", "rmdf_envmap_bind_cubes", "rmdf_envmap_get_cube", "rmdf_envmap_unbind",
    "rmdf_set_cornell", "rmdf_cornell_box_vertices", "rmdf_render_output_pixels", "rmdf_render", "rmdf_render_device",
    "rmdf_julia", "rmdf_julia_device", "rmdf_julia_bands_device", "rmdf_mip_resolve_device", "rmdf_counters_get", "rmdf_counters_reset", "rmdf_set_work_order", "rmdf_measure_fp32_peak", "rmdf_measure_mufu_peak",
    "rmdf_shared_alloc", "rmdf_shared_free", "rmdf_shared_open", "rmdf_shared_close", "rmdf_enable_peer",
    "rmdf_host_register", "rmdf_host_unregister",
    "rmdf_frame_begin", "rmdf_frame_end", "rmdf_frame_sync_status",
]


class Params(C.Structure):
    """rmdf_params"""
    _fields_ = [
        ("scene", C.c_int32), ("fb_w", C.c_int32), ("fb_h", C.c_int32), ("time", C.c_float),
        ("tile_idx", C.c_int32), ("resolve_log2", C.c_int32), ("out_format", C.c_int32), ("mode", C.c_int32),
        ("lod_policy", C.c_int32), ("shard_rank", C.c_int32), ("shard_count", C.c_int32), ("layout", C.c_int32),
        ("shard_grid_x", C.c_int32), ("shard_grid_y", C.c_int32), ("shard_skew", C.c_int32), ("flags", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32),
    ]


class Counters(C.Structure):
    """rmdf_counters"""
    _fields_ = [(n, C.c_uint64) for n in ("samples", "in_sphere", "hits", "march_de", "shade_de", "inner_iters", "launches")] + \
               [("last_kernel_ms", C.c_float), ("reserved", C.c_float)]

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n in ("samples", "in_sphere", "hits", "march_de", "shade_de", "inner_iters", "launches")}
        d["last_kernel_ms"] = float(self.last_kernel_ms)
        return d


class RmdfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("rmdf error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load(build_if_missing: bool = True):
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise ImportError("librmdf_b200.so is not built (run python build.py); there is no CPU fallback")
        from . import build as _b
        _b.build()
    L = C.CDLL(str(LIB_PATH))
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)
    PP = C.POINTER(Params)
    L.rmdf_create.argtypes = [C.POINTER(vp), C.c_int]
    L.rmdf_destroy.argtypes = [vp]; L.rmdf_destroy.restype = None
    L.rmdf_last_error.argtypes = []; L.rmdf_last_error.restype = C.c_char_p
    L.rmdf_abi_version.argtypes = []
    L.rmdf_ctx_last_error.argtypes = [vp, C.c_char_p, C.c_int]
    L.rmdf_set_stream.argtypes = [vp, vp]
    L.rmdf_sync.argtypes = [vp]
    L.rmdf_envmap_resize_height.argtypes = [C.c_int] * 3
    L.rmdf_envmap_resize.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
    L.rmdf_envmap_convolve.argtypes = [vp, vp, C.c_int, C.c_int, fp, C.c_int, vp]
    L.rmdf_envmap_convolve_device.argtypes = [vp, vp, C.c_int, C.c_int, fp, C.c_int, vp]
    L.rmdf_envmap_convolve_rows_device.argtypes = [vp, vp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int, vp]
    L.rmdf_rgbe_roundtrip.argtypes = [vp, C.c_int64]
    L.rmdf_envmap_bind.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
    L.rmdf_envmap_bind_cube.argtypes = [vp, C.c_int, vp, C.c_int]
    L.rmdf_envmap_bind_cubes.argtypes = [vp, C.c_int, ip, C.POINTER(vp), ip]
    L.rmdf_envmap_get_cube.argtypes = [vp, C.c_int, vp, ip]
    L.rmdf_envmap_unbind.argtypes = [vp, C.c_int]
    L.rmdf_set_cornell.argtypes = [vp, vp, C.c_int]
    L.rmdf_cornell_box_vertices.argtypes = [vp]
    L.rmdf_render_output_pixels.argtypes = [PP, C.POINTER(C.c_int32)]; L.rmdf_render_output_pixels.restype = C.c_int64
    L.rmdf_render.argtypes = [vp, PP, vp]
    L.rmdf_render_device.argtypes = [vp, PP, vp]
    L.rmdf_julia.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp]
    L.rmdf_julia_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp]
    L.rmdf_julia_bands_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.rmdf_mip_resolve_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
    L.rmdf_counters_get.argtypes = [vp, C.POINTER(Counters)]
    L.rmdf_counters_reset.argtypes = [vp]
    L.rmdf_set_work_order.argtypes = [vp, C.c_int]
    L.rmdf_measure_fp32_peak.argtypes = [vp, fp, fp]
    L.rmdf_measure_mufu_peak.argtypes = [vp, fp]
    L.rmdf_shared_alloc.argtypes = [vp, C.c_int64, C.POINTER(vp), C.c_char_p]
    L.rmdf_shared_free.argtypes = [vp, vp]
    L.rmdf_shared_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.rmdf_shared_close.argtypes = [vp, vp]
    L.rmdf_enable_peer.argtypes = [vp, C.c_int]
    L.rmdf_host_register.argtypes = [vp, vp, C.c_int64, C.POINTER(vp)]
    L.rmdf_host_unregister.argtypes = [vp, vp]
    L.rmdf_frame_begin.argtypes = [vp, vp, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int]
    L.rmdf_frame_end.argtypes = [vp, vp, C.c_int, C.c_uint32, C.c_int]
    L.rmdf_frame_sync_status.argtypes = [vp, ip]
    for name in EXPORTS:
        f = getattr(L, name)
        if name not in ("rmdf_destroy", "rmdf_last_error", "rmdf_render_output_pixels"):
            f.restype = C.c_int
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise RmdfError(rc, load().rmdf_last_error().decode("utf-8", "replace"))
