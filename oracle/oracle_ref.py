"""ctypes loader for oracle/_ref/libref_fragment.so — the reference's OWN fragment shader
(/root/reference/fragment.shd) compiled by g++ through oracle/ref/glsl_compat.hpp.

TEST INFRASTRUCTURE ONLY (same rule as oracle.py: tests/, smoke() and bench.py's CPU legs).
Built in the dev container by `make -C oracle ref` (__graft_entry__.build() does it when
/root/reference exists); the GPU box has no reference tree and uses the prebuilt .so that
travels with the snapshot.  `available()` says whether it can be used.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from . import oracle as O

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_ref" / "libref_fragment.so"
REFERENCE = Path("/root/reference")

# variant index → (scene, oracle flags); 0-3 = FragmentShader enum (ShaderRendering.hs:46-47,119-123)
VARIANTS = {0: (O.SCENE_CORNELL, 0), 1: (O.SCENE_TEST, 0), 2: (O.SCENE_MB_POWER8, 0), 3: (O.SCENE_MB_GENERAL, 0),
            4: (O.SCENE_CORNELL, O.FLAG_RAY_TRACING_TEST), 5: (O.SCENE_TEST, O.FLAG_CAM_ORTHO),
            6: (O.SCENE_MB_POWER8, O.FLAG_CAM_ORTHO)}


def variant_for(scene: int, flags: int = 0) -> int:
    for v, (s, f) in VARIANTS.items():
        if s == scene and f == flags:
            return v
    raise KeyError("no compiled shader variant for scene %d flags %d" % (scene, flags))


def build(force: bool = False) -> Path | None:
    """Compile from the reference tree when it is present; otherwise keep the prebuilt file."""
    O.build()
    shader = REFERENCE / "fragment.shd"
    if shader.exists():
        srcs = [shader, _HERE / "ref" / "glsl_compat.hpp", _HERE / "ref" / "ref_driver.cpp", _HERE / "ref" / "gen_shader.py",
                _HERE / "oracle.h", _HERE.parent / "include" / "rmdf_glsl.h", _HERE / "Makefile", O._LIB_PATH]
        stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
        if force or stale:
            r = subprocess.run(["make", "-C", str(_HERE), "ref"] + (["-B"] if force else []),
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode:
                raise RuntimeError("oracle/_ref build failed:\n" + r.stdout)
    return _LIB_PATH if _LIB_PATH.exists() else None


def available() -> bool:
    try:
        return build() is not None
    except RuntimeError:
        return False


_lib = None


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/libref_fragment.so is missing and /root/reference is not here to build it")
        O.lib()                                   # liboracle.so first (the sampler lives there)
        L = C.CDLL(str(_LIB_PATH))
        fp, u8p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
        i64 = C.c_int64
        for v in VARIANTS:
            s = "_v%d" % v
            getattr(L, "ref_render" + s).argtypes = [C.POINTER(O._Frame), C.c_int, C.c_int, C.c_int, C.c_int, u32p, fp, u8p, C.POINTER(C.c_uint64)]
            getattr(L, "ref_render" + s).restype = C.c_int
            getattr(L, "ref_distance_estimator" + s).argtypes = [C.c_float, fp, fp, fp, i64]
            getattr(L, "ref_ray_march" + s).argtypes = [C.c_float, fp, fp, fp, fp, i64]
            getattr(L, "ref_normal" + s).argtypes = [C.c_float, fp, C.c_int, fp, fp, i64]
            getattr(L, "ref_distance_ao" + s).argtypes = [C.c_float, fp, fp, fp, fp, i64]
            getattr(L, "ref_fresnel_conductor" + s).argtypes = [fp, C.c_float, C.c_float, fp, i64]
            getattr(L, "ref_lookat" + s).argtypes = [fp, fp, fp, fp]
            getattr(L, "ref_generate_ray" + s).argtypes = [fp, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, fp]
            getattr(L, "ref_ray_sphere" + s).argtypes = [fp, fp, C.c_float, fp, i64]
            getattr(L, "ref_triplex_pow8" + s).argtypes = [fp, fp, i64]
            getattr(L, "ref_triplex_pow" + s).argtypes = [fp, C.c_float, fp, i64]
            getattr(L, "ref_de_triangle" + s).argtypes = [fp, fp, fp, i64]
            getattr(L, "ref_smin" + s).argtypes = [fp, fp, C.c_float, fp, i64]
        _lib = L
    return _lib


_fp, _f32 = O._fp, O._f32


def _fn(name: str, variant: int):
    return getattr(lib(), "%s_v%d" % (name, variant))


def render(scene: int, fb_w: int, fb_h: int, time: float = 0.0, envs=None, rect=None,
           lod_policy: int = O.LOD_QUAD, threads: int = 0, want_linear: bool = True, flags: int = 0):
    """Same contract as oracle.render(), executed by the reference's shader text."""
    v = variant_for(scene, flags)
    x0, y0, x1, y1 = rect if rect is not None else (0, 0, fb_w, fb_h)
    f = O._Frame()
    f.scene, f.fb_w, f.fb_h, f.time, f.lod_policy, f.threads, f.flags = scene, fb_w, fb_h, time, lod_policy, threads, flags
    keep = []
    for s in range(5):
        e = None if envs is None else envs[s]
        if e is None:
            f.env_face[s] = 0
        else:
            e = _f32(e); keep.append(e)
            f.env_face[s] = e.shape[1]
            f.env[s] = _fp(e)
    cv = O.cornell_vertices()
    f.cornell = _fp(cv)
    w, h = x1 - x0, y1 - y0
    fb = np.empty((h, w), np.uint32)
    lin = np.empty((h, w, 3), np.float32) if want_linear else None
    hit = np.empty((h, w), np.uint8)
    nq = C.c_uint64()
    rc = _fn("ref_render", v)(C.byref(f), x0, y0, x1, y1, fb.ctypes.data_as(C.POINTER(C.c_uint32)),
                              _fp(lin) if lin is not None else None, hit.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nq))
    if rc != 0:
        raise ValueError("ref_render failed with code %d" % rc)
    return {"fb": fb, "linear": lin, "hit": hit, "nearest_quads": int(nq.value)}


def distance(variant: int, time: float, pts) -> np.ndarray:
    pts = _f32(pts).reshape(-1, 3); cv = O.cornell_vertices(); out = np.empty(len(pts), np.float32)
    _fn("ref_distance_estimator", variant)(C.c_float(time), _fp(cv), _fp(pts), _fp(out), len(pts))
    return out


def ray_march(variant: int, time: float, origins, dirs) -> np.ndarray:
    o = _f32(origins).reshape(-1, 3); d = _f32(dirs).reshape(-1, 3); cv = O.cornell_vertices()
    out = np.empty((len(o), 2), np.float32)
    _fn("ref_ray_march", variant)(C.c_float(time), _fp(cv), _fp(o), _fp(d), _fp(out), len(o))
    return out


def normal(variant: int, time: float, pts, central: bool = False) -> np.ndarray:
    pts = _f32(pts).reshape(-1, 3); cv = O.cornell_vertices(); out = np.empty_like(pts)
    _fn("ref_normal", variant)(C.c_float(time), _fp(cv), int(central), _fp(pts), _fp(out), len(pts))
    return out


def distance_ao(variant: int, time: float, pts, normals) -> np.ndarray:
    pts = _f32(pts).reshape(-1, 3); nn = _f32(normals).reshape(-1, 3); cv = O.cornell_vertices()
    out = np.empty(len(pts), np.float32)
    _fn("ref_distance_ao", variant)(C.c_float(time), _fp(cv), _fp(pts), _fp(nn), _fp(out), len(pts))
    return out


def fresnel_conductor(cosi, eta: float, k: float, variant: int = 1) -> np.ndarray:
    c = _f32(cosi).ravel(); out = np.empty_like(c)
    _fn("ref_fresnel_conductor", variant)(_fp(c), C.c_float(eta), C.c_float(k), _fp(out), c.size)
    return out


def lookat(eye, focus=(0, 0, 0), up=(0, 1, 0), variant: int = 1) -> np.ndarray:
    e, f, u = _f32(eye), _f32(focus), _f32(up); out = np.empty(12, np.float32)
    _fn("ref_lookat", variant)(_fp(e), _fp(f), _fp(u), _fp(out))
    return out


def generate_ray(cam, w: float, h: float, px: int, py: int, ortho: bool = False, width_or_hfov: float = 67.5, variant: int = 1) -> np.ndarray:
    c = _f32(cam); out = np.empty(6, np.float32)
    _fn("ref_generate_ray", variant)(_fp(c), C.c_float(w), C.c_float(h), px, py, int(ortho), C.c_float(width_or_hfov), _fp(out))
    return out


def ray_sphere(origins, dirs, radius: float, variant: int = 1) -> np.ndarray:
    o = _f32(origins).reshape(-1, 3); d = _f32(dirs).reshape(-1, 3); out = np.empty((len(o), 3), np.float32)
    _fn("ref_ray_sphere", variant)(_fp(o), _fp(d), C.c_float(radius), _fp(out), len(o))
    return out


def triplex_pow8(pts, variant: int = 2) -> np.ndarray:
    p = _f32(pts).reshape(-1, 3); out = np.empty_like(p)
    _fn("ref_triplex_pow8", variant)(_fp(p), _fp(out), len(p))
    return out


def triplex_pow(pts, power: float, variant: int = 3) -> np.ndarray:
    p = _f32(pts).reshape(-1, 3); out = np.empty_like(p)
    _fn("ref_triplex_pow", variant)(_fp(p), C.c_float(power), _fp(out), len(p))
    return out


def de_triangle(pts, tri, variant: int = 0) -> np.ndarray:
    p = _f32(pts).reshape(-1, 3); t = _f32(tri).ravel(); out = np.empty(len(p), np.float32)
    _fn("ref_de_triangle", variant)(_fp(p), _fp(t), _fp(out), len(p))
    return out


def smin(a, b, k: float, variant: int = 1) -> np.ndarray:
    a = _f32(a).ravel(); b = _f32(b).ravel(); out = np.empty_like(a)
    _fn("ref_smin", variant)(_fp(a), _fp(b), C.c_float(k), _fp(out), a.size)
    return out
