"""ctypes loader for the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; the product package never
imports this module.  PARITY STATUS: the shader half is pinned by oracle/_ref (oracle_ref.py,
oracle_march.c); the Haskell half (oracle_envmap.c, oracle_julia.c) is "parity unpinned".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"

SCENE_CORNELL, SCENE_TEST, SCENE_MB_POWER8, SCENE_MB_GENERAL = 0, 1, 2, 3
ENV_REFLECTION, ENV_COS_1, ENV_COS_8, ENV_COS_64, ENV_COS_512 = range(5)
LOD_QUAD, LOD_ALWAYS_LINEAR = 0, 1
FLAG_NORMAL_CENTRAL, FLAG_CAM_ORTHO, FLAG_RAY_TRACING_TEST = 1, 2, 4


class _Frame(C.Structure):
    _fields_ = [
        ("scene", C.c_int32), ("fb_w", C.c_int32), ("fb_h", C.c_int32), ("time", C.c_float),
        ("lod_policy", C.c_int32), ("threads", C.c_int32), ("flags", C.c_int32),
        ("env_face", C.c_int32 * 5), ("env", C.POINTER(C.c_float) * 5),
        ("cornell", C.POINTER(C.c_float)),
    ]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("samples", "in_sphere", "hits", "march_de", "shade_de", "mb_iters", "nan_pixels", "nearest_quads")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    srcs = list(_HERE.glob("*.c")) + [_HERE / "oracle.h", _HERE.parent / "include" / "rmdf_glsl.h"]
    stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        try:
            subprocess.run(["make", "-C", str(_HERE)] + (["-B"] if force else []), check=True,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        except (subprocess.CalledProcessError, FileNotFoundError) as e:  # pragma: no cover
            if not _LIB_PATH.exists():
                raise RuntimeError("oracle build failed: %s" % getattr(e, "stdout", e))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        fp, u8p, u32p, ip = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_int)
        L.oracle_render.argtypes = [C.POINTER(_Frame), C.c_int, C.c_int, C.c_int, C.c_int, u32p, fp, u8p, C.POINTER(Counters)]
        L.oracle_render.restype = C.c_int
        L.oracle_camera.argtypes = [C.c_int, C.c_float, fp]
        L.oracle_mb_power.argtypes = [C.c_float]; L.oracle_mb_power.restype = C.c_float
        L.oracle_distance.argtypes = [C.c_int, C.c_float, fp, C.c_float, C.c_float, C.c_float]
        L.oracle_distance.restype = C.c_float
        L.oracle_glsl_eval.argtypes = [C.c_int, fp, fp, fp, C.c_int64]
        i64 = C.c_int64
        L.oracle_distance_n.argtypes = [C.c_int, C.c_float, fp, fp, fp, i64]
        L.oracle_ray_march_n.argtypes = [C.c_int, C.c_float, fp, fp, fp, fp, i64]
        L.oracle_normal_n.argtypes = [C.c_int, C.c_float, fp, C.c_int, fp, fp, i64]
        L.oracle_distance_ao_n.argtypes = [C.c_int, C.c_float, fp, fp, fp, fp, i64]
        L.oracle_fresnel_conductor_n.argtypes = [fp, C.c_float, C.c_float, fp, i64]
        L.oracle_generate_ray.argtypes = [fp, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, fp]
        L.oracle_ray_sphere_n.argtypes = [fp, fp, C.c_float, fp, i64]
        L.oracle_triplex_pow_n.argtypes = [fp, C.c_float, C.c_int, fp, i64]
        L.oracle_de_triangle_n.argtypes = [fp, fp, fp, i64]
        L.oracle_smin_n.argtypes = [fp, fp, C.c_float, fp, i64]
        L.oracle_hdr_decode.argtypes = [u8p, C.c_int64, ip, ip, fp, C.c_int64]; L.oracle_hdr_decode.restype = C.c_int
        L.oracle_hdr_encode.argtypes = [fp, C.c_int, C.c_int, u8p, C.c_int64]; L.oracle_hdr_encode.restype = C.c_int64
        L.oracle_rgbe_roundtrip.argtypes = [fp, C.c_int64, fp]
        L.oracle_resize_height.argtypes = [C.c_int] * 3; L.oracle_resize_height.restype = C.c_int
        L.oracle_resize.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
        L.oracle_convolve.argtypes = [fp, C.c_int, C.c_int, C.c_float, fp, C.c_int]
        L.oracle_convolve_rows.argtypes = [fp, C.c_int, C.c_int, C.c_float, fp, C.c_int, C.c_int, C.c_int]
        L.oracle_convolve_sweep_rows.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, fp, C.c_int, C.c_int]
        L.oracle_julia_segments.argtypes = [C.c_int, C.c_int, u32p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
        L.oracle_latlong_to_cube.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int]
        L.oracle_cube_pixel_to_dir.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, fp]
        L.oracle_test_latlong_envmap.argtypes = [fp]
        L.oracle_cornell_vertices.argtypes = [fp]
        L.oracle_make_segments.argtypes = [C.c_int, C.c_int, C.c_int, ip, C.c_int]; L.oracle_make_segments.restype = C.c_int
        L.oracle_julia.argtypes = [C.c_int, C.c_int, u32p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
        L.oracle_mip_resolve.argtypes = [u32p, C.c_int, C.c_int, C.c_int, u32p]
        L.oracle_screenshot_rows.argtypes = [u32p, C.c_int, C.c_int, u8p]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------- march / shade
def cornell_vertices() -> np.ndarray:
    out = np.empty((96, 3), np.float32)
    lib().oracle_cornell_vertices(_fp(out))
    return out


def camera(scene: int, time: float) -> np.ndarray:
    out = np.empty(12, np.float32)
    lib().oracle_camera(scene, C.c_float(time), _fp(out))
    return out


def distance(scene: int, time: float, pts: np.ndarray) -> np.ndarray:
    pts = _f32(pts).reshape(-1, 3)
    cv = cornell_vertices()
    out = np.empty(len(pts), np.float32)
    lib().oracle_distance_n(scene, C.c_float(time), _fp(cv), _fp(pts), _fp(out), len(pts))
    return out


# single shader functions over arrays (lattice tests against oracle_ref)
def ray_march(scene: int, time: float, origins, dirs) -> np.ndarray:
    o = _f32(origins).reshape(-1, 3); d = _f32(dirs).reshape(-1, 3); cv = cornell_vertices()
    out = np.empty((len(o), 2), np.float32)
    lib().oracle_ray_march_n(scene, C.c_float(time), _fp(cv), _fp(o), _fp(d), _fp(out), len(o))
    return out


def normal(scene: int, time: float, pts, central: bool = False) -> np.ndarray:
    pts = _f32(pts).reshape(-1, 3); cv = cornell_vertices(); out = np.empty_like(pts)
    lib().oracle_normal_n(scene, C.c_float(time), _fp(cv), int(central), _fp(pts), _fp(out), len(pts))
    return out


def distance_ao(scene: int, time: float, pts, normals) -> np.ndarray:
    pts = _f32(pts).reshape(-1, 3); nn = _f32(normals).reshape(-1, 3); cv = cornell_vertices()
    out = np.empty(len(pts), np.float32)
    lib().oracle_distance_ao_n(scene, C.c_float(time), _fp(cv), _fp(pts), _fp(nn), _fp(out), len(pts))
    return out


def fresnel_conductor(cosi, eta: float, k: float) -> np.ndarray:
    c = _f32(cosi).ravel(); out = np.empty_like(c)
    lib().oracle_fresnel_conductor_n(_fp(c), C.c_float(eta), C.c_float(k), _fp(out), c.size)
    return out


def generate_ray(cam, w: float, h: float, px: int, py: int, ortho: bool = False) -> np.ndarray:
    c = _f32(cam); out = np.empty(6, np.float32)
    lib().oracle_generate_ray(_fp(c), C.c_float(w), C.c_float(h), px, py, int(ortho), _fp(out))
    return out


def ray_sphere(origins, dirs, radius: float) -> np.ndarray:
    o = _f32(origins).reshape(-1, 3); d = _f32(dirs).reshape(-1, 3); out = np.empty((len(o), 3), np.float32)
    lib().oracle_ray_sphere_n(_fp(o), _fp(d), C.c_float(radius), _fp(out), len(o))
    return out


def triplex_pow(pts, power: float = 8.0, power8: bool = True) -> np.ndarray:
    p = _f32(pts).reshape(-1, 3); out = np.empty_like(p)
    lib().oracle_triplex_pow_n(_fp(p), C.c_float(power), int(power8), _fp(out), len(p))
    return out


def de_triangle(pts, tri) -> np.ndarray:
    p = _f32(pts).reshape(-1, 3); t = _f32(tri).ravel(); out = np.empty(len(p), np.float32)
    lib().oracle_de_triangle_n(_fp(p), _fp(t), _fp(out), len(p))
    return out


def smin(a, b, k: float) -> np.ndarray:
    a = _f32(a).ravel(); b = _f32(b).ravel(); out = np.empty_like(a)
    lib().oracle_smin_n(_fp(a), _fp(b), C.c_float(k), _fp(out), a.size)
    return out


def render(scene: int, fb_w: int, fb_h: int, time: float = 0.0, envs=None, rect=None,
           lod_policy: int = LOD_QUAD, threads: int = 0, want_linear: bool = True, flags: int = 0):
    """fragment.shd main() for every pixel of rect = (x0, y0, x1, y1) (default: full frame).

    envs: None (all slots constant 1.0 = "no env lighting") or a list of 5
    arrays [6, w, w, 3] float32 (RGB16F values) / None per slot, in texture-unit order.
    Returns dict(fb uint32 [h, w] row 0 = bottom, linear float32 [h, w, 3], hit uint8 [h, w],
    counters dict).
    """
    x0, y0, x1, y1 = rect if rect is not None else (0, 0, fb_w, fb_h)
    f = _Frame()
    f.scene, f.fb_w, f.fb_h, f.time, f.lod_policy, f.threads, f.flags = scene, fb_w, fb_h, time, lod_policy, threads, flags
    keep = []
    for s in range(5):
        e = None if envs is None else envs[s]
        if e is None:
            f.env_face[s] = 0
        else:
            e = _f32(e)
            assert e.ndim == 4 and e.shape[0] == 6 and e.shape[1] == e.shape[2] and e.shape[3] == 3
            keep.append(e)
            f.env_face[s] = e.shape[1]
            f.env[s] = _fp(e)
    cv = cornell_vertices()
    f.cornell = _fp(cv)
    w, h = x1 - x0, y1 - y0
    fb = np.empty((h, w), np.uint32)
    lin = np.empty((h, w, 3), np.float32) if want_linear else None
    hit = np.empty((h, w), np.uint8)
    cnt = Counters()
    rc = lib().oracle_render(C.byref(f), x0, y0, x1, y1, fb.ctypes.data_as(C.POINTER(C.c_uint32)),
                             _fp(lin) if lin is not None else None, hit.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(cnt))
    if rc != 0:
        raise ValueError("oracle_render failed with code %d" % rc)
    return {"fb": fb, "linear": lin, "hit": hit, "counters": cnt.as_dict()}


GLSL_FN = {"log": 0, "exp": 1, "pow": 2, "sin": 3, "cos": 4, "tan": 5, "acos": 6, "atan2": 7,
           "inversesqrt": 8, "round_to_half": 9, "unorm8": 10, "mod": 11}


def glsl(fn: str, x, y=None) -> np.ndarray:
    """Element-wise include/rmdf_glsl.h built-in `fn` (float32 in, float32 out)."""
    x = _f32(x).ravel()
    y = _f32(np.zeros_like(x) if y is None else np.broadcast_to(y, x.shape)).ravel()
    out = np.empty_like(x)
    lib().oracle_glsl_eval(GLSL_FN[fn], _fp(x), _fp(y), _fp(out), x.size)
    return out


# ---------------------------------------------------------------- env maps
def hdr_decode(data: bytes) -> np.ndarray:
    buf = np.frombuffer(data, np.uint8)
    w, h = C.c_int(), C.c_int()
    p = buf.ctypes.data_as(C.POINTER(C.c_uint8))
    rc = lib().oracle_hdr_decode(p, len(data), C.byref(w), C.byref(h), None, 0)
    if rc:
        raise ValueError("not a Radiance RGBE image (code %d)" % rc)
    out = np.empty((h.value, w.value, 3), np.float32)
    rc = lib().oracle_hdr_decode(p, len(data), C.byref(w), C.byref(h), _fp(out), out.size)
    if rc:
        raise ValueError("truncated Radiance image (code %d)" % rc)
    return out


def hdr_encode(img: np.ndarray) -> bytes:
    img = _f32(img)
    h, w, _ = img.shape
    n = lib().oracle_hdr_encode(_fp(img), w, h, None, 0)
    out = np.empty(n, np.uint8)
    lib().oracle_hdr_encode(_fp(img), w, h, out.ctypes.data_as(C.POINTER(C.c_uint8)), n)
    return out.tobytes()


def rgbe_roundtrip(img: np.ndarray) -> np.ndarray:
    img = _f32(img)
    out = np.empty_like(img)
    lib().oracle_rgbe_roundtrip(_fp(img), img.size // 3, _fp(out))
    return out


def resize(img: np.ndarray, dstw: int) -> np.ndarray:
    img = _f32(img)
    h, w, _ = img.shape
    dsth = lib().oracle_resize_height(w, h, dstw)
    out = np.empty((dsth, dstw, 3), np.float32)
    lib().oracle_resize(_fp(img), w, h, dstw, _fp(out))
    return out


def convolve(img: np.ndarray, power: float, threads: int = 0) -> np.ndarray:
    img = _f32(img)
    h, w, _ = img.shape
    out = np.empty_like(img)
    lib().oracle_convolve(_fp(img), w, h, C.c_float(power), _fp(out), threads)
    return out


def convolve_rows(img: np.ndarray, power: float, rows, threads: int = 0) -> np.ndarray:
    """destination rows [y0, y1) only; returns those rows"""
    img = _f32(img)
    h, w, _ = img.shape
    out = np.zeros_like(img)
    lib().oracle_convolve_rows(_fp(img), w, h, C.c_float(power), _fp(out), rows[0], rows[1], threads)
    return out[rows[0]:rows[1]]


def convolve_sweep_rows(img: np.ndarray, powers, rows) -> np.ndarray:
    """The reference's decomposition (ShaderRendering.hs:142): one thread per power, each sequential."""
    img = _f32(img)
    h, w, _ = img.shape
    pw = _f32(list(powers))
    out = np.zeros((len(pw),) + img.shape, np.float32)
    lib().oracle_convolve_sweep_rows(_fp(img), w, h, _fp(pw), len(pw), _fp(out), rows[0], rows[1])
    return out[:, rows[0]:rows[1]]


def latlong_to_cube(img: np.ndarray, round_half: bool = True) -> np.ndarray:
    img = _f32(img)
    h, w, _ = img.shape
    fw = w // 3
    out = np.empty((6, fw, fw, 3), np.float32)
    lib().oracle_latlong_to_cube(_fp(img), w, h, _fp(out), int(round_half))
    return out


def cube_pixel_to_dir(face: int, w: int, x: int, y: int) -> np.ndarray:
    out = np.empty(3, np.float32)
    lib().oracle_cube_pixel_to_dir(face, w, x, y, _fp(out))
    return out


def test_latlong_envmap() -> np.ndarray:
    out = np.empty((256, 512, 3), np.float32)
    lib().oracle_test_latlong_envmap(_fp(out))
    return out


def make_segments(nseg: int, low: int, high: int):
    cap = max(1, min(nseg, max(high - low, 0)) + 1)
    buf = (C.c_int * (2 * cap))()
    n = lib().oracle_make_segments(nseg, low, high, buf, cap)
    return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]


# ---------------------------------------------------------------- julia / frame buffer
def julia(w: int, h: int, smooth: bool, tick: float, rows=None, threads: int = 0, segments: int = 0) -> np.ndarray:
    """segments > 0: the reference's forSegmentsConcurrently decomposition (Fractal2D.hs:98) — the rows cut
    by makeNSegments into `segments` contiguous bands, one thread each; else OpenMP dynamic over rows."""
    y0, y1 = rows if rows is not None else (0, h)
    fb = np.zeros((y1 - y0, w), np.uint32)
    # the C function indexes px + py·w with absolute py: hand it the address row 0 would have
    base = C.cast(C.c_void_p(fb.ctypes.data - y0 * w * 4), C.POINTER(C.c_uint32))
    if segments > 0:
        lib().oracle_julia_segments(w, h, base, int(smooth), C.c_double(tick), y0, y1, segments)
    else:
        lib().oracle_julia(w, h, base, int(smooth), C.c_double(tick), y0, y1, threads)
    return fb


def mip_resolve(fb: np.ndarray, levels: int) -> np.ndarray:
    fb = np.ascontiguousarray(fb, np.uint32)
    h, w = fb.shape
    out = np.empty((h >> levels, w >> levels), np.uint32)
    lib().oracle_mip_resolve(fb.ctypes.data_as(C.POINTER(C.c_uint32)), w, h, levels, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


def screenshot_rows(fb: np.ndarray) -> np.ndarray:
    fb = np.ascontiguousarray(fb, np.uint32)
    h, w = fb.shape
    out = np.empty((h, w, 4), np.uint8)
    lib().oracle_screenshot_rows(fb.ctypes.data_as(C.POINTER(C.c_uint32)), w, h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def host_threads() -> int:
    """The threads the CPU legs really use: the oracle's `threads` argument is always passed explicitly
    (num_threads clause), so OMP_NUM_THREADS — which torchrun sets to 1 — does not apply."""
    try:
        return len(os.sched_getaffinity(0)) or 1
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1
