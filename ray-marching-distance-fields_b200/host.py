"""Host-side mirror of the reference's interface for the hot path, on top of the C ABI.

The reference's host is Haskell (ShaderRendering.hs, FrameBuffer.hs, HDREnvMap.hs,
Fractal2D.hs); GHC is not in this image, so the Haskell shim lives uncompiled in
haskell/RMDF/FFI.hs and this module mirrors the same names, argument meaning and
error behaviour in Python so that the parity tests read like calls into the
reference.  Everything computational goes through librmdf_b200.so — there is no CPU
path here (file I/O and the RGBE byte format excepted).

    reference                                   this module
    ---------------------------------------------------------------------------
    withShaderRenderer        (ShaderRendering.hs:60)   withShaderRenderer / ShaderRenderer
    drawShaderTile            (ShaderRendering.hs:151)  ShaderRenderer.drawShaderTile
    FragmentShader(..)        (ShaderRendering.hs:46)   FragmentShader
    isTileIdxFirst/LastTile   (ShaderRendering.hs:54)   isTileIdxFirstTile / isTileIdxLastTile
    withFrameBuffer & co.     (FrameBuffer.hs:51-228)   FrameBuffer
    loadHDRImage              (HDREnvMap.hs:31)         loadHDRImage
    resizeHDRImage            (HDREnvMap.hs:169)        ShaderRenderer.resizeHDRImage
    cosineConvolveHDREnvMap   (HDREnvMap.hs:217)        ShaderRenderer.cosineConvolveHDREnvMap
    latLongHDREnvMapToCubeMap (HDREnvMap.hs:118)        ShaderRenderer.latLongHDREnvMapToCubeMap
    buildPreConvolvedHDREnvMapCache (ShaderRendering.hs:131)  ShaderRenderer.buildPreConvolvedHDREnvMapCache
    juliaAnimated             (Fractal2D.hs:63)         ShaderRenderer.juliaAnimated
"""
from __future__ import annotations

import contextlib
import ctypes as C
import enum
import os
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import Params, Counters, RmdfError, check


class FragmentShader(enum.IntEnum):
    """ShaderRendering.hs:46-47 (Enum order = C ABI scene id)."""
    FSDECornellBoxShader = 0
    FSDETestShader = 1
    FSMBPower8Shader = 2
    FSMBGeneralShader = 3


tilesX, tilesY = 8, 8
nTiles = tilesX * tilesY


def isTileIdxLastTile(idx: int) -> bool:
    return idx % nTiles == nTiles - 1


def isTileIdxFirstTile(idx: int) -> bool:
    return idx % nTiles == 0


# ---------------------------------------------------------------- Radiance .hdr I/O
def _rgbe_to_float(rgbe: np.ndarray) -> np.ndarray:
    """JuicyPixels toFloat: (mantissa + 0.0)·2^(e-136)."""
    f = np.ldexp(np.float32(1.0), rgbe[..., 3].astype(np.int32) - 136).astype(np.float32)
    return (rgbe[..., :3].astype(np.float32) * f[..., None]).astype(np.float32)


def _float_to_rgbe(img: np.ndarray) -> np.ndarray:
    """JuicyPixels toRGBE."""
    img = np.ascontiguousarray(img, np.float32)
    d = img.max(axis=-1)
    sig, e = np.frexp(d)
    with np.errstate(divide="ignore", invalid="ignore"):
        coeff = (sig.astype(np.float32) * np.float32(255.9999) / d).astype(np.float32)
    m = np.trunc(img * coeff[..., None]).astype(np.int64) & 255
    out = np.empty(img.shape[:-1] + (4,), np.uint8)
    out[..., :3] = m
    out[..., 3] = (e + 128).astype(np.uint8)
    out[d <= np.float32(1e-32)] = 0
    return out


def decodeHDR(data: bytes) -> np.ndarray:
    """Radiance RGBE (flat or new-style RLE) → float32 [h, w, 3], row 0 = top."""
    pos = 0
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end]
        pos = end + 1
        if len(line) == 0:
            break
    end = data.index(b"\n", pos)
    res = data[pos:end].split()
    pos = end + 1
    if len(res) != 4 or res[0] != b"-Y" or res[2] != b"+X":
        raise ValueError("unsupported Radiance resolution line %r" % (data[pos:end],))
    h, w = int(res[1]), int(res[3])
    buf = np.frombuffer(data, np.uint8)
    if len(data) - pos == w * h * 4:
        rgbe = buf[pos:pos + w * h * 4].reshape(h, w, 4)
        return _rgbe_to_float(rgbe)
    rgbe = np.empty((h, w, 4), np.uint8)
    for y in range(h):
        if not (buf[pos] == 2 and buf[pos + 1] == 2 and ((int(buf[pos + 2]) << 8) | int(buf[pos + 3])) == w):
            rgbe[y] = buf[pos:pos + w * 4].reshape(w, 4)
            pos += w * 4
            continue
        pos += 4
        for c in range(4):
            x = 0
            while x < w:
                cnt = int(buf[pos]); pos += 1
                if cnt > 128:
                    cnt -= 128
                    rgbe[y, x:x + cnt, c] = buf[pos]; pos += 1
                else:
                    rgbe[y, x:x + cnt, c] = buf[pos:pos + cnt]; pos += cnt
                x += cnt
    return _rgbe_to_float(rgbe)


def encodeHDR(img: np.ndarray) -> bytes:
    h, w, _ = img.shape
    hdr = ("#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (h, w)).encode("ascii")
    return hdr + _float_to_rgbe(img).tobytes()


def loadHDRImage(fn) -> np.ndarray:
    """HDREnvMap.hs:31-52.  Raises ValueError (the reference's `Left err`)."""
    try:
        with open(fn, "rb") as f:
            return decodeHDR(f.read())
    except (OSError, ValueError, IndexError) as e:
        raise ValueError("Not an HDR RGBF image: %s (%s)" % (fn, e))


def saveRadianceImage(fn, img: np.ndarray) -> None:
    """JP.saveRadianceImage with the reference's cleanup (ShaderRendering.hs:146-148)."""
    try:
        with open(fn, "wb") as f:
            f.write(encodeHDR(img))
    except BaseException:
        with contextlib.suppress(OSError):
            os.remove(fn)
        raise


# ---------------------------------------------------------------- frame buffer
class FrameBuffer:
    """FrameBuffer.hs:39-228 without GL: an RGBA8 buffer of w·h Word32 (0xAABBGGRR,
    idx = px + py·w, row 0 = bottom), optional supersampling and screenshots."""

    def __init__(self, w: int, h: int, downscaling: str = "HighQualityDownscaling"):
        self.fbDownscaling = downscaling
        self.resizeFrameBuffer(w, h)

    def resizeFrameBuffer(self, w: int, h: int) -> None:
        self.fbDim = (int(w), int(h))
        self.data = np.zeros((int(h), int(w)), np.uint32)
        self.data[:] = 0xFF000000          # cleared to opaque black (FrameBuffer.hs:109-111)

    def getFrameBufferDim(self):
        return self.fbDim

    def fillFrameBuffer(self, f):
        """f(w, h, vec) writes the whole buffer; vec is the flat Word32 view (:117-158)."""
        w, h = self.fbDim
        return f(w, h, self.data.reshape(-1))

    def drawIntoFrameBuffer(self, f):
        w, h = self.fbDim
        return f(w, h)

    def saveFrameBufferToPNG(self, fn) -> None:
        """Level 0, rows flipped, alpha forced to 0xFF (:215-228)."""
        from PIL import Image
        rgba = self.data[::-1].copy().view(np.uint8).reshape(self.fbDim[1], self.fbDim[0], 4)
        rgba[..., 3] = 0xFF
        Image.fromarray(rgba, "RGBA").save(fn)


# ---------------------------------------------------------------- renderer
def _as_f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class ShaderRenderer:
    """Owns one rmdf_ctx (one GPU).  Mirrors the ShaderRenderer record
    (ShaderRendering.hs:36-44): env cube maps on texture units 0-4 + Cornell geometry."""

    powers = (1.0, 8.0, 64.0, 512.0)          # ShaderRendering.hs:71

    def __init__(self, device: int = 0):
        self._lib = _capi.load()
        h = C.c_void_p()
        check(self._lib.rmdf_create(C.byref(h), device))
        self._h = h
        self.device = device
        self.srEnvCubeMaps = []
        xyz = np.empty((96, 3), np.float32)
        check(self._lib.rmdf_cornell_box_vertices(xyz.ctypes.data))
        self.cornellBoxVertices = xyz
        check(self._lib.rmdf_set_cornell(self._h, xyz.ctypes.data, 96))     # mkCornellBoxVerticesTex

    # -- lifetime ----------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.rmdf_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):  # pragma: no cover
        with contextlib.suppress(Exception):
            self.close()

    @property
    def handle(self):
        return self._h

    def setStream(self, cuda_stream: Optional[int]):
        check(self._lib.rmdf_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        check(self._lib.rmdf_sync(self._h))

    def setWorkOrder(self, enabled: bool):
        """rmdf_set_work_order: claim work items most-expensive-first, from the previous frame's item costs (scheduling only)."""
        check(self._lib.rmdf_set_work_order(self._h, 1 if enabled else 0))

    # -- HDREnvMap.hs ---------------------------------------------------------------
    def resizeHDRImage(self, img, dstw: int) -> np.ndarray:
        img = _as_f32(img)
        h, w, _ = img.shape
        dsth = self._lib.rmdf_envmap_resize_height(w, h, dstw)
        out = np.empty((dsth, dstw, 3), np.float32)
        check(self._lib.rmdf_envmap_resize(self._h, img.ctypes.data, w, h, dstw, out.ctypes.data))
        return out

    def cosineConvolveHDREnvMapSweep(self, img, powers: Sequence[float]) -> np.ndarray:
        """All powers in one pass → [np, h, w, 3]."""
        img = _as_f32(img)
        h, w, _ = img.shape
        p = (C.c_float * len(powers))(*[float(x) for x in powers])
        out = np.empty((len(powers), h, w, 3), np.float32)
        check(self._lib.rmdf_envmap_convolve(self._h, img.ctypes.data, w, h, p, len(powers), out.ctypes.data))
        return out

    def cosineConvolveHDREnvMap(self, img, power: float) -> np.ndarray:
        return self.cosineConvolveHDREnvMapSweep(img, [power])[0]

    @staticmethod
    def rgbeRoundTrip(img) -> np.ndarray:
        out = _as_f32(img).copy()
        check(_capi.load().rmdf_rgbe_roundtrip(out.ctypes.data, out.size // 3))
        return out

    def latLongHDREnvMapToCubeMap(self, slot: int, latlong) -> None:
        latlong = _as_f32(latlong)
        h, w, _ = latlong.shape
        check(self._lib.rmdf_envmap_bind(self._h, slot, latlong.ctypes.data, w, h))

    def bindCubeMap(self, slot: int, faces) -> None:
        faces = _as_f32(faces)
        assert faces.ndim == 4 and faces.shape[0] == 6 and faces.shape[1] == faces.shape[2] and faces.shape[3] == 3
        check(self._lib.rmdf_envmap_bind_cube(self._h, slot, faces.ctypes.data, faces.shape[1]))

    def bindCubeMaps(self, slots, faces_list) -> None:
        """Several slots in one batch (one synchronisation): per-frame uploads from pinned host memory.
        faces_list: numpy arrays or torch CPU tensors [6, w, w, 3] float32."""
        n = len(faces_list)
        ptrs = (C.c_void_p * n)(); ws = (C.c_int * n)(); sl = (C.c_int * n)(*[int(x) for x in slots])
        keep = []
        for i, f in enumerate(faces_list):
            if hasattr(f, "data_ptr"):
                ptrs[i] = f.data_ptr(); ws[i] = f.shape[1]
            else:
                f = _as_f32(f); keep.append(f)
                ptrs[i] = f.ctypes.data; ws[i] = f.shape[1]
        check(self._lib.rmdf_envmap_bind_cubes(self._h, n, sl, ptrs, ws))

    def unbindCubeMap(self, slot: int) -> None:
        check(self._lib.rmdf_envmap_unbind(self._h, slot))

    def getCubeMap(self, slot: int) -> Optional[np.ndarray]:
        fw = C.c_int(0)
        check(self._lib.rmdf_envmap_get_cube(self._h, slot, None, C.byref(fw)))
        if fw.value == 0:
            return None
        out = np.empty((6, fw.value, fw.value, 3), np.float32)
        check(self._lib.rmdf_envmap_get_cube(self._h, slot, out.ctypes.data, C.byref(fw)))
        return out

    # -- ShaderRendering.hs:60-110,131-149 -----------------------------------------------
    @staticmethod
    def mkCacheFn(reflMapFn, power: float) -> str:
        """`dropExtension reflMapFn ++ "_cache_pow_" ++ show pow ++ ".hdr"` (show 8.0 = "8.0")."""
        return os.path.splitext(str(reflMapFn))[0] + "_cache_pow_" + repr(float(power)) + ".hdr"

    def buildPreConvolvedHDREnvMapCache(self, reflMap, powfn) -> int:
        """Computes the missing cache files; returns how many were written."""
        missing = [(p, fn) for p, fn in powfn if not os.path.exists(fn)]
        if not missing:
            return 0
        resized = self.resizeHDRImage(reflMap, 256)                     # :139
        conv = self.cosineConvolveHDREnvMapSweep(resized, [p for p, _ in missing])
        for (p, fn), img in zip(missing, conv):
            saveRadianceImage(fn, img)
        return len(missing)

    def loadEnvironment(self, reflMapFn, cache_dir=None) -> None:
        """The env-map half of withShaderRenderer (:66-91): load, build/verify the cache,
        re-load every cache file, convert to cube maps, bind to texture units 0-4."""
        reflMap = loadHDRImage(reflMapFn)
        base = Path(reflMapFn) if cache_dir is None else Path(cache_dir) / Path(reflMapFn).name
        powfn = [(p, self.mkCacheFn(base, p)) for p in self.powers]
        self.buildPreConvolvedHDREnvMapCache(reflMap, powfn)
        self.srEnvCubeMaps = ["env_reflection"]
        self.latLongHDREnvMapToCubeMap(_capi.ENV_REFLECTION, reflMap)
        for i, (p, fn) in enumerate(powfn):
            self.latLongHDREnvMapToCubeMap(_capi.ENV_COS_1 + i, loadHDRImage(fn))
            self.srEnvCubeMaps.append("env_cos_%d" % int(round(p)))

    # -- ShaderRendering.hs:151-196 ----------------------------------------------------
    def makeParams(self, shdEnum, tileIdx: Optional[int], w: int, h: int, time: float, *, resolve_log2=0,
                   out_format=_capi.OUT_RGBA8, mode=_capi.MODE_PARITY, lod_policy=_capi.LOD_QUAD,
                   shard_rank=0, shard_count=1, layout=_capi.LAYOUT_FRAME, shard_grid=(0, 0), shard_skew=0, flags=0, rows=(0, 0)) -> Params:
        p = Params()
        p.scene = int(shdEnum); p.fb_w = int(w); p.fb_h = int(h)
        p.time = float(np.float32(time))                 # realToFrac :: Double -> Float (:171)
        p.tile_idx = -1 if tileIdx is None else int(tileIdx)
        p.resolve_log2 = resolve_log2; p.out_format = out_format; p.mode = mode; p.lod_policy = lod_policy
        p.shard_rank = shard_rank; p.shard_count = shard_count; p.layout = layout
        p.shard_grid_x, p.shard_grid_y = shard_grid; p.shard_skew = shard_skew; p.flags = flags
        p.row_begin, p.row_end = rows
        return p

    def outputPixels(self, p: Params):
        nt = C.c_int32(0)
        n = self._lib.rmdf_render_output_pixels(C.byref(p), C.byref(nt))
        if n < 0:
            check(_capi.ERR_ARG)
        return int(n), int(nt.value)

    def drawShaderTile(self, shdEnum, tileIdx: Optional[int], w: int, h: int, time: float, fb=None, **kw) -> np.ndarray:
        """Renders tile `tileIdx` (None = whole frame) of a w×h frame buffer into `fb`
        (uint32 [h, w] row 0 = bottom, or float32 [h, w, 4] for OUT_RGBA32F).  With
        resolve_log2 = k the result is mip level k ([h>>k, w>>k])."""
        p = self.makeParams(shdEnum, tileIdx, w, h, time, **kw)
        n, _ = self.outputPixels(p)
        k = p.resolve_log2
        if p.layout == _capi.LAYOUT_FRAME:
            shape = (h >> k, w >> k)
        else:
            shape = (n,)
        if p.out_format == _capi.OUT_RGBA32F:
            shape, dtype = shape + (4,), np.float32
        else:
            dtype = np.uint32
        if fb is None:
            fb = np.zeros(shape, dtype)
        assert fb.dtype == dtype and fb.size == int(np.prod(shape)) and fb.flags["C_CONTIGUOUS"]
        check(self._lib.rmdf_render(self._h, C.byref(p), fb.ctypes.data))
        return fb

    def renderDevice(self, p: Params, dev_ptr: int) -> None:
        """Asynchronous render into device memory (e.g. a torch CUDA tensor's data_ptr())."""
        check(self._lib.rmdf_render_device(self._h, C.byref(p), C.c_void_p(dev_ptr)))

    def renderHost(self, p: Params, host_ptr: int) -> None:
        check(self._lib.rmdf_render(self._h, C.byref(p), C.c_void_p(host_ptr)))

    # -- Fractal2D.hs:63-98 --------------------------------------------------------------
    def juliaAnimated(self, w: int, h: int, fb, smooth: bool, tick: float, rows=None, resolve_log2: int = 0) -> np.ndarray:
        y0, y1 = rows if rows is not None else (0, h)
        shape = (max(y1 - y0, 0) >> resolve_log2, max(w, 0) >> resolve_log2)
        if fb is None:
            fb = np.zeros(shape, np.uint32)
        assert fb.dtype == np.uint32 and fb.size == shape[0] * shape[1]
        check(self._lib.rmdf_julia(self._h, w, h, int(bool(smooth)), C.c_double(tick), y0, y1, resolve_log2, fb.ctypes.data))
        return fb

    def juliaDevice(self, w, h, smooth, tick, rows, resolve_log2, dev_ptr: int) -> None:
        y0, y1 = rows
        check(self._lib.rmdf_julia_device(self._h, w, h, int(bool(smooth)), C.c_double(tick), y0, y1, resolve_log2, C.c_void_p(dev_ptr)))

    def juliaBandsDevice(self, w, h, smooth, tick, resolve_log2, band_h, shard_rank, shard_count, dev_ptr: int,
                         layout=_capi.LAYOUT_PACKED_TILES) -> None:
        check(self._lib.rmdf_julia_bands_device(self._h, w, h, int(bool(smooth)), C.c_double(tick), resolve_log2, band_h,
                                                shard_rank, shard_count, layout, C.c_void_p(dev_ptr)))

    # -- multi-GPU shared frame buffer ------------------------------------------------------
    def sharedAlloc(self, nbytes: int):
        """(device pointer, 64-byte IPC handle) of a buffer other processes' kernels can write into."""
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        check(self._lib.rmdf_shared_alloc(self._h, nbytes, C.byref(ptr), handle))
        return int(ptr.value), handle.raw

    def sharedFree(self, ptr: int) -> None:
        check(self._lib.rmdf_shared_free(self._h, C.c_void_p(ptr)))

    def sharedOpen(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        check(self._lib.rmdf_shared_open(self._h, handle, C.byref(ptr)))
        return int(ptr.value)

    def sharedClose(self, ptr: int) -> None:
        check(self._lib.rmdf_shared_close(self._h, C.c_void_p(ptr)))

    def frameBegin(self, ctrl_ptr: int, rank: int, world: int, seq: int, timeout_ms: int = 10000, deal: int = _capi.DEAL_INTERLEAVED) -> None:
        """One frame over several GPUs (include/rmdf.h rmdf_frame_begin): arms this context's next render."""
        check(self._lib.rmdf_frame_begin(self._h, C.c_void_p(ctrl_ptr), rank, world, seq, deal, timeout_ms))

    def frameEnd(self, ctrl_ptr: int, world: int, seq: int, timeout_ms: int = 10000) -> None:
        check(self._lib.rmdf_frame_end(self._h, C.c_void_p(ctrl_ptr), world, seq, timeout_ms))

    def frameSyncStatus(self) -> int:
        """0, or 1 + the index of the control word a wait gave up on (synchronises the stream)."""
        v = C.c_int(0)
        check(self._lib.rmdf_frame_sync_status(self._h, C.byref(v)))
        return int(v.value)

    def enablePeer(self, peer_device: int) -> None:
        check(self._lib.rmdf_enable_peer(self._h, peer_device))

    def hostRegister(self, host_ptr: int, nbytes: int) -> int:
        """Page-lock host memory for this context's device; returns the device-side address of it."""
        ptr = C.c_void_p()
        check(self._lib.rmdf_host_register(self._h, C.c_void_p(host_ptr), nbytes, C.byref(ptr)))
        return int(ptr.value)

    def hostUnregister(self, host_ptr: int) -> None:
        check(self._lib.rmdf_host_unregister(self._h, C.c_void_p(host_ptr)))

    # -- counters ------------------------------------------------------------------------
    def counters(self) -> dict:
        c = Counters()
        check(self._lib.rmdf_counters_get(self._h, C.byref(c)))
        return c.as_dict()

    def measureFp32Peak(self) -> dict:
        t, f = C.c_float(0), C.c_float(0)
        check(self._lib.rmdf_measure_fp32_peak(self._h, C.byref(t), C.byref(f)))
        return {"tflops": float(t.value), "sm_mhz": float(f.value)}

    def measureMufuPeak(self) -> float:
        """1e9 MUFU lane-operations per second (the XU-pipe roofline denominator)."""
        g = C.c_float(0)
        check(self._lib.rmdf_measure_mufu_peak(self._h, C.byref(g)))
        return float(g.value)

    def resetCounters(self) -> None:
        check(self._lib.rmdf_counters_reset(self._h))


@contextlib.contextmanager
def withShaderRenderer(reflMapFn=None, device: int = 0, cache_dir=None):
    """Bracket-style lifetime like ShaderRendering.hs:60.  The shader file argument of
    the reference has no counterpart: the CUDA scenes are compiled ahead of time."""
    sr = ShaderRenderer(device)
    try:
        if reflMapFn is not None:
            sr.loadEnvironment(reflMapFn, cache_dir)
        yield sr
    finally:
        sr.close()
