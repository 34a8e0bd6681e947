"""CPU tests of the oracle itself (no GPU): the committed golden vectors, an independent
numpy/float32 re-derivation of the hot formulas straight from the reference text, the
reference's own debug fixtures (buildTestLatLongEnvMap, HDREnvMap.hs:55-73) and the edge
cases of its helpers (makeNSegments, ConcurrentSegments.hs:14-23).

The reference ships no tests and cannot be executed here, so these pin the
oracle to (a) its frozen known answers and (b) a second, independently written restatement."""
import json
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
f32 = np.float32


# ------------------------------------------------------------------ golden vectors
def test_golden_de_lattice(O):
    g = json.loads((GOLD / "de_lattice.json").read_text())
    pts = np.array(g["points"], np.uint32).view(np.float32)
    for s in range(4):
        got = O.distance(s, g["time"], pts).view(np.uint32)
        assert np.array_equal(got, np.array(g["de_bits"][str(s)], np.uint32)), "scene %d" % s


def _synthetic_envs(O):
    rng = np.random.RandomState(1234)
    envs = []
    for w in (6, 4, 4, 3, 3):
        e = rng.uniform(0.0, 2.0, size=(6, w, w, 3)).astype(np.float32)
        envs.append(O.glsl("round_to_half", e).reshape(e.shape))
    return envs


def test_golden_frames(O):
    z = np.load(GOLD / "frames_64.npz")
    envs = _synthetic_envs(O)
    for s in range(4):
        for t in (0.0, 7.0):
            assert np.array_equal(O.render(s, 64, 64, t)["fb"], z["white_s%d_t%d" % (s, int(t))])
            assert np.array_equal(O.render(s, 64, 64, t, envs=envs)["fb"], z["env_s%d_t%d" % (s, int(t))])


def test_ref_golden_reference_shader_outputs(O):
    """tests/golden/ref_shader.npz holds outputs of the reference's OWN fragment shader (compiled by g++
    from /root/reference/fragment.shd, tests/golden/make_ref_golden.py): frames per shader variant (RGBA8
    words, linear colour bits, hit masks), the commented-out shader switches, crops of BASELINE configs
    1/3/4 at their real sizes and a distance-estimator lattice.  The C oracle must reproduce all of it."""
    z = np.load(GOLD / "ref_shader.npz")
    envs = _synthetic_envs(O)
    for s in range(4):
        for t in (0.0, 7.0):
            for name, e in (("white", None), ("env", envs)):
                r = O.render(s, 64, 48, t, envs=e)
                k = "%s_s%d_t%d" % (name, s, int(t))
                assert np.array_equal(r["fb"], z["fb_" + k]), k
                lin, want = r["linear"].view(np.uint32), z["lin_" + k]
                nan = np.isnan(r["linear"])
                assert np.array_equal(nan, np.isnan(want.view(np.float32))) and np.array_equal(lin[~nan], want[~nan]), k
                assert np.array_equal(np.packbits(r["hit"]), z["hit_" + k]), k
    for s, fl in ((0, O.FLAG_RAY_TRACING_TEST), (1, O.FLAG_CAM_ORTHO), (2, O.FLAG_CAM_ORTHO)):
        assert np.array_equal(O.render(s, 64, 48, 1.0, envs=envs, flags=fl)["fb"], z["fb_flag%d_s%d" % (fl, s)])
    assert np.array_equal(O.render(2, 512, 512, 0.0, rect=(224, 240, 288, 272))["fb"], z["fb_c1_crop"])
    assert np.array_equal(O.render(0, 3840, 2160, 0.0, rect=(1888, 1064, 1952, 1096))["fb"], z["fb_c3_cornell_crop"])
    assert np.array_equal(O.render(1, 3840, 2160, 0.0, rect=(1888, 1064, 1952, 1096))["fb"], z["fb_c3_csg_crop"])
    assert np.array_equal(O.render(2, 15360, 8640, 0.0, rect=(7648, 4304, 7712, 4336))["fb"], z["fb_c4_crop"])
    pts = z["de_points"].view(np.float32)
    for s in range(4):
        got, want = O.distance(s, 7.0, pts), z["de_s%d" % s].view(np.float32)
        nan = np.isnan(got)
        assert np.array_equal(nan, np.isnan(want)) and np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32)), s


def test_golden_julia_camera_cornell(O):
    z = np.load(GOLD / "julia_96x64.npz")
    for t in (0.0, 7.0, 123.5):
        for sm in (0, 1):
            assert np.array_equal(O.julia(96, 64, bool(sm), t), z["j_t%d_s%d" % (int(t), sm)])
    cams = json.loads((GOLD / "camera.json").read_text())
    for key, bits in cams.items():
        s, t = key.split("_")
        assert np.array_equal(O.camera(int(s), float(t)).view(np.uint32), np.array(bits, np.uint32))
    assert np.array_equal(O.cornell_vertices(), np.load(GOLD / "cornell_vertices.npy"))


def test_golden_test_envmap_convolution(O):
    z = np.load(GOLD / "test_envmap_conv.npz")
    small = O.resize(O.test_latlong_envmap(), 32)
    assert np.array_equal(small, z["small"])
    assert np.array_equal(O.convolve(small, 1.0), z["p1"])
    assert np.array_equal(O.convolve(small, 8.0, threads=1), z["p8"])       # thread count must not matter


# ------------------------------------------------------------------ independent restatements
def _np_mandelbulb8(O, p):
    """fragment.shd:74-99,101-158 (POWER8) written again in numpy float32, vectorised."""
    p = np.asarray(p, f32)
    pos = np.stack([p[:, 2], p[:, 0], p[:, 1]], 1)                 # pos.zxy
    w = pos.copy()
    dr = np.ones(len(p), f32)
    r = np.zeros(len(p), f32)
    live = np.ones(len(p), bool)
    with np.errstate(all="ignore"):
        for _ in range(25):
            rr = np.sqrt((w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1]) + w[:, 2] * w[:, 2]).astype(f32)
            r = np.where(live, rr, r)
            live = live & ~(r > f32(4))
            x, y, z = w[:, 0], w[:, 1], w[:, 2]
            x2 = x * x; x4 = x2 * x2; y2 = y * y; y4 = y2 * y2; z2 = z * z; z4 = z2 * z2
            k3 = y2 + x2
            k2 = (f32(1) / np.sqrt(k3 * k3 * k3 * k3 * k3 * k3 * k3)).astype(f32)
            k1 = y4 + z4 + x4 - f32(6) * z2 * x2 - f32(6) * y2 * z2 + f32(2) * x2 * y2
            k4 = y2 - z2 + x2
            nx = f32(-8) * z * k4 * (y4 * y4 - f32(28) * y4 * y2 * x2 + f32(70) * y4 * x4 - f32(28) * y2 * x2 * x4 + x4 * x4) * k1 * k2
            ny = f32(64) * y * z * x * (y2 - x2) * k4 * (y4 - f32(6) * y2 * x2 + x4) * k1 * k2
            nz = f32(-16) * z2 * k3 * k4 * k4 + k1 * k1
            nw = np.stack([nx, ny, nz], 1).astype(f32) + pos
            r2 = r * r; r4 = r2 * r2
            ndr = ((r4 * r2) * r) * f32(8) * dr + f32(1)
            w = np.where(live[:, None], nw, w)
            dr = np.where(live, ndr, dr).astype(f32)
    return (f32(0.5) * O.glsl("log", r) * r / dr).astype(f32)


def test_mandelbulb8_matches_numpy_restatement(O):
    rng = np.random.RandomState(7)
    pts = rng.uniform(-1.2, 1.2, size=(4000, 3)).astype(f32)
    ref = _np_mandelbulb8(O, pts)
    got = O.distance(O.SCENE_MB_POWER8, 0.0, pts)
    ok = np.isfinite(ref)
    assert ok.mean() > 0.99
    assert np.array_equal(got[ok].view(np.uint32), ref[ok].view(np.uint32))


def test_csg_scene_matches_numpy_restatement(O):
    """fragment.shd:447-456 with libm exp/log in float64 — the deterministic built-ins are within
    a few ulp, so values agree to ~1e-6."""
    rng = np.random.RandomState(8)
    p = rng.uniform(-1, 1, size=(3000, 3))
    ln = lambda a, b: np.sqrt(a * a + b * b)
    smin = lambda a, b, k: -np.log(np.exp(-k * a) + np.exp(-k * b)) / k
    torus = lambda q: ln(ln(q[:, 0], q[:, 1]) - 0.85, q[:, 2]) - 0.1
    box = lambda q, b: np.sqrt((np.maximum(np.abs(q) - np.array(b), 0) ** 2).sum(1)) - 0.03
    d_sphere = np.sqrt((p * p).sum(1)) - 0.4
    d_torus = smin(smin(torus(p), torus(p[:, [2, 0, 1]]), 64), torus(p[:, [1, 2, 0]]), 64)
    d_box = smin(smin(box(p, (0.8, 0.06, 0.06)), box(p, (0.06, 0.8, 0.06)), 64), box(p, (0.06, 0.06, 0.8)), 64)
    ref = smin(d_box, np.minimum(d_sphere, d_torus), 64)
    got = O.distance(O.SCENE_TEST, 0.0, p.astype(f32))
    assert np.allclose(got, ref, rtol=0, atol=3e-6)


def test_cornell_de_matches_bruteforce(O):
    """Cornell DE (fragment.shd:312-411) against a float64 closest-point-on-triangle search."""
    v = O.cornell_vertices().astype(np.float64).reshape(32, 3, 3)
    rng = np.random.RandomState(9)
    pts = rng.uniform(-0.9, 0.9, size=(300, 3))

    def tri_dist(p, a, b, c):
        # Ericson, Real-Time Collision Detection §5.1.5
        ab, ac, ap = b - a, c - a, p - a
        d1, d2 = ab @ ap, ac @ ap
        if d1 <= 0 and d2 <= 0: return np.linalg.norm(p - a)
        bp = p - b; d3, d4 = ab @ bp, ac @ bp
        if d3 >= 0 and d4 <= d3: return np.linalg.norm(p - b)
        vc = d1 * d4 - d3 * d2
        if vc <= 0 and d1 >= 0 and d3 <= 0: return np.linalg.norm(p - (a + ab * (d1 / (d1 - d3))))
        cp = p - c; d5, d6 = ab @ cp, ac @ cp
        if d6 >= 0 and d5 <= d6: return np.linalg.norm(p - c)
        vb = d5 * d2 - d1 * d6
        if vb <= 0 and d2 >= 0 and d6 <= 0: return np.linalg.norm(p - (a + ac * (d2 / (d2 - d6))))
        va = d3 * d6 - d5 * d4
        if va <= 0 and (d4 - d3) >= 0 and (d5 - d6) >= 0:
            return np.linalg.norm(p - (b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)))))
        den = 1.0 / (va + vb + vc)
        return np.linalg.norm(p - (a + ab * (vb * den) + ac * (vc * den)))

    ref = np.array([min(tri_dist(p, *t) for t in v) for p in pts])
    got = O.distance(O.SCENE_CORNELL, 0.0, pts.astype(f32))
    assert np.allclose(got, ref, atol=2e-5)


def test_julia_matches_numpy_restatement(O):
    """Fractal2D.hs:63-98 again in numpy float32 (non-smooth: integers must agree exactly)."""
    w, h, tick = 80, 48, 11.25
    ft = f32(tick)
    frac = lambda q: q - np.trunc(q)
    a, b, c = frac(ft / f32(17)), frac(ft / f32(61)), frac(ft / f32(71))
    two_pi = a * f32(2) * f32(np.pi)
    jr = f32(np.sin(two_pi, dtype=f32)) * max(f32(0.7), b)
    ji = f32(np.cos(two_pi, dtype=f32)) * max(f32(0.7), c)
    fw, fh = f32(w), f32(h)
    ratio = fw / fh
    py, px = np.meshgrid(np.arange(h, dtype=f32), np.arange(w, dtype=f32), indexing="ij")
    y = (py / fh) * f32(2.9) - f32(1.45)
    x = (px / fw) * f32(2.9) * ratio - f32(1.45) * ratio
    zr, zi = x.copy(), y.copy()
    it = np.zeros((h, w), np.int32)
    live = np.ones((h, w), bool)
    for _ in range(40):
        esc = (zr * zr + zi * zi) > f32(16)
        live = live & ~esc
        nr = (zr * zr - zi * zi) + jr
        ni = (zr * zi + zi * zr) + ji
        cyc = live & (nr == zr) & (ni == zi)
        it = np.where(cyc, 40, it)
        live = live & ~cyc
        zr = np.where(live, nr, zr); zi = np.where(live, ni, zi)
        it = np.where(live, it + 1, it)
    ref = (np.trunc(it.astype(f32) / f32(40) * f32(255)).astype(np.uint32) << 8)
    assert np.array_equal(O.julia(w, h, False, tick), ref)
    sm = O.julia(w, h, True, tick)
    assert (sm & 0xFFFF00FF == 0).all() and (sm >> 8 <= 255).all()
    assert np.array_equal(sm[it == 40], ref[it == 40])          # interior = maxIter in both variants


# ------------------------------------------------------------------ reference helpers
def test_make_segments_edge_cases(O):
    """makeNSegments, ConcurrentSegments.hs:14-23."""
    assert O.make_segments(4, 5, 5) == []                        # low >= high
    assert O.make_segments(0, 0, 10) == []                       # nsegc <= 0
    assert O.make_segments(1, 3, 9) == [(3, 9)]
    assert O.make_segments(3, 0, 10) == [(0, 3), (3, 6), (6, 10)]   # last takes the remainder
    assert O.make_segments(8, 0, 3) == [(0, 1), (1, 2), (2, 3)]     # min(N, range) segments
    assert O.make_segments(4, 2, 10) == [(2, 4), (4, 6), (6, 8), (8, 10)]
    for n, lo, hi in [(7, 0, 100), (16, 3, 1080), (5, 0, 5)]:
        segs = O.make_segments(n, lo, hi)
        assert segs[0][0] == lo and segs[-1][1] == hi and all(a[1] == b[0] for a, b in zip(segs, segs[1:]))


def test_cube_pixel_to_dir_is_inverse_of_gl_face_selection(O):
    """cubeMapPixelToDir (HDREnvMap.hs:76-87) vs the GL face table the sampler uses: sampling a cube
    map at a texel-centre direction with NEAREST returns exactly that texel."""
    w = 5
    faces = np.arange(6 * w * w * 3, dtype=np.float32).reshape(6, w, w, 3)
    faces = O.glsl("round_to_half", faces).reshape(faces.shape)
    # render() cannot sample arbitrary directions, so check through the background: a 1×1 frame
    # looks down the camera axis; instead verify the geometric inverse directly
    for face in range(6):
        for (x, y) in [(0, 0), (w - 1, 0), (2, 3), (w - 1, w - 1)]:
            d = O.cube_pixel_to_dir(face, w, x, y)
            ax = np.abs(d)
            major = int(np.argmax(ax))
            sel = major * 2 + (1 if d[major] < 0 else 0)
            assert sel == face
            sc, tc, ma = {0: (-d[2], -d[1], d[0]), 1: (d[2], -d[1], -d[0]), 2: (d[0], d[2], d[1]),
                          3: (d[0], -d[2], -d[1]), 4: (d[0], -d[1], d[2]), 5: (-d[0], -d[1], -d[2])}[face]
            i, j = int(np.floor(0.5 * (sc / ma + 1) * w)), int(np.floor(0.5 * (tc / ma + 1) * w))
            assert (i, j) == (x, y)


def test_test_latlong_envmap_and_cube(O):
    """buildTestLatLongEnvMap (HDREnvMap.hs:55-73): six flat colours; every cube face of the
    converted map is (away from the seams) its debug colour (:149-156 uses the same table)."""
    env = O.test_latlong_envmap()
    assert env.shape == (256, 512, 3)
    cols = {tuple(c) for c in env.reshape(-1, 3)}
    assert cols == {(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 0), (0, 1, 1)}
    cube = O.latlong_to_cube(env)
    assert cube.shape == (6, 170, 170, 3)
    expect = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 0)]   # +X -X +Y -Y +Z(back) -Z(front)
    for f in range(6):
        centre = cube[f, 40:130, 40:130].reshape(-1, 3)
        assert (centre == np.array(expect[f], np.float32)).all(), f


def test_uffizi_fixture_and_rgbe(O, uffizi):
    assert uffizi.shape == (256, 512, 3)
    assert uffizi.max() == np.float32(15.9375)                     # 255·2^-4 (JuicyPixels: mantissa + 0.0)
    assert np.allclose(uffizi.mean(axis=(0, 1)), [1.7846, 1.7548, 2.0330], atol=1e-3)
    # toRGBE truncates v·(significand·255.9999/d) while toFloat adds 0.0: a write/read cycle of
    # already-quantised data lands at most one mantissa step below (never above)
    back = O.hdr_decode(O.hdr_encode(uffizi))
    step = np.ldexp(np.float32(1), np.frexp(uffizi.max(axis=-1))[1] - 8)[..., None]
    assert (back <= uffizi).all() and (uffizi - back <= step).all()
    assert np.array_equal(back, O.rgbe_roundtrip(uffizi))
    rt = O.rgbe_roundtrip(uffizi * np.float32(0.37))
    rel = np.abs(rt - uffizi * np.float32(0.37)).max(axis=-1) / (uffizi * np.float32(0.37)).max(axis=-1)
    assert rel.max() < 1.0 / 128.0                                 # 8-bit mantissa, truncated
    assert O.rgbe_roundtrip(np.zeros((2, 3), f32)).sum() == 0


def test_resize_and_convolve_properties(O, uffizi):
    small = O.resize(uffizi, 256)
    assert small.shape == (128, 256, 3)
    assert np.allclose(small.mean(axis=(0, 1)), uffizi.mean(axis=(0, 1)), rtol=2e-3)
    tiny = O.resize(uffizi, 32)
    one = np.ones_like(tiny)
    c1 = O.convolve(one, 1.0)
    # constant map: Σ sinθ·c / n — linear in the input, same for every channel
    assert np.allclose(c1[..., 0], c1[..., 1]) and np.allclose(O.convolve(one * 3, 1.0), c1 * 3, rtol=1e-5)
    # higher powers are darker (cos^p <= cos for 0 <= cos <= 1)
    assert (O.convolve(tiny, 8.0).sum() < O.convolve(tiny, 1.0).sum())


def test_mip_resolve_and_screenshot(O):
    fb = np.array([[0xFF000000, 0xFF0000FF], [0xFF00FF00, 0xFFFF0003]], np.uint32)
    out = O.mip_resolve(fb, 1)
    # per byte (a+b+c+d+2)>>2 : R (0+255+0+3+2)>>2=65, G (0+0+255+0+2)>>2=64, B 64, A 255
    assert out.shape == (1, 1) and out[0, 0] == (0xFF << 24) | (64 << 16) | (64 << 8) | 65
    rng = np.random.RandomState(3)
    big = rng.randint(0, 2 ** 32, size=(16, 24), dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(O.mip_resolve(big, 2), O.mip_resolve(O.mip_resolve(big, 1), 1))
    shot = O.screenshot_rows(fb)
    assert shot.shape == (2, 2, 4) and (shot[..., 3] == 255).all()
    assert tuple(shot[0, 0][:3]) == (0, 255, 0)                    # PNG row 0 = texture row h-1


def test_render_rect_equals_crop_of_full_frame(O):
    full = O.render(2, 96, 64, 0.0)
    part = O.render(2, 96, 64, 0.0, rect=(32, 16, 80, 48))
    assert np.array_equal(part["fb"], full["fb"][16:48, 32:80])
    assert np.array_equal(part["hit"], full["hit"][16:48, 32:80])


def test_camera_matches_reference_constants(O):
    """in_time = 0: eye (0, 1.7071, 1.7071) for the orbit camera, (0, 0.4, -2) for Cornell."""
    cam = O.camera(2, 0.0)
    assert np.allclose(cam[9:], [0.0, 1.70710678, 1.70710678], atol=1e-6)
    assert np.allclose(O.camera(0, 0.0)[9:], [0.0, 0.4, -2.0], atol=1e-7)
    x, y, z = cam[0:3], cam[3:6], cam[6:9]
    assert abs(np.dot(x, y)) < 1e-6 and abs(np.dot(y, z)) < 1e-6 and abs(np.linalg.norm(z) - 1) < 1e-6


def test_ray_traced_cornell_cross_checks_the_distance_estimator(O):
    """RAY_TRACING_TEST (fragment.shd:915-937) is the reference's own cross-check of the Cornell DE:
    the sphere tracer's hit mask must match the ray tracer's, up to the
    silhouette pixels where MIN_DIST = 0.001 fattens the surface."""
    w, h = 320, 180
    de = O.render(O.SCENE_CORNELL, w, h, 0.0)["hit"].astype(bool)
    rt = O.render(O.SCENE_CORNELL, w, h, 0.0, flags=O.FLAG_RAY_TRACING_TEST)
    agree = (de == rt["hit"].astype(bool)).mean()
    assert agree > 0.985, agree
    assert (de | ~rt["hit"].astype(bool)).mean() > 0.999          # every ray-traced hit is a sphere-traced hit
    # hit colours are (n + 1) / 2 of axis-aligned walls: only a handful of distinct colours
    cols = {tuple(c) for c in rt["fb"][rt["hit"].astype(bool)].reshape(-1)[:, None].view(np.uint8)[:, :3]}
    assert 3 <= len(cols) <= 40


def test_option_flags_change_only_what_they_say(O):
    base = O.render(2, 96, 96, 0.0)
    cen = O.render(2, 96, 96, 0.0, flags=O.FLAG_NORMAL_CENTRAL)
    assert np.array_equal(base["hit"], cen["hit"])                 # normals do not move the surface
    assert cen["counters"]["shade_de"] == base["counters"]["hits"] * 8       # 6 normal + 2 AO evaluations per hit
    assert base["counters"]["shade_de"] == base["counters"]["hits"] * 6
    ortho = O.render(2, 96, 96, 0.0, flags=O.FLAG_CAM_ORTHO)
    assert ortho["hit"].mean() > base["hit"].mean()                # width-2 ortho view frames the bulb tighter


# ------------------------------------------------------------------ whole-pipeline second restatement
def _np_render_mandelbulb_white(O, w, h, t):
    """fragment.shd main() for the POWER8 Mandelbulb with all env maps = 1.0, written a second time,
    vectorised in numpy float32 straight from the shader text (camera :884-902, lookat :829-838,
    generate_ray :840-871, ray_sphere :595-616, ray_march :618-676, normal :463-470, distance_ao
    :542-562, fresnel :694-719, shading :799-810, gamma :956-960).  Uses only IEEE float32 numpy
    arithmetic plus the deterministic built-ins (O.glsl) for sin/cos/tan/log/pow."""
    g1 = lambda fn, x, y=None: O.glsl(fn, np.asarray([x], f32), None if y is None else np.asarray([y], f32))[0]
    T = f32(t)
    cp = np.array([g1("sin", T / f32(3)), g1("cos", T / f32(4)), g1("cos", T / f32(3))], f32)
    dot = lambda a, b: (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]
    inv = lambda d: (f32(1) / np.sqrt(d)).astype(f32)
    norm = lambda v: (v * inv(dot(v, v))[..., None]).astype(f32)
    eye = (norm(cp) * f32(2.414213562373095)).astype(f32)
    z = norm(eye - f32(0)); up = np.array([0, 1, 0], f32)
    cross = lambda a, b: np.stack([a[..., 1] * b[..., 2] - b[..., 1] * a[..., 2], a[..., 2] * b[..., 0] - b[..., 2] * a[..., 0],
                                   a[..., 0] * b[..., 1] - b[..., 0] * a[..., 1]], -1).astype(f32)
    x = norm(cross(up, z)); y = cross(z, x)
    py, px = np.meshgrid(np.arange(h, dtype=f32), np.arange(w, dtype=f32), indexing="ij")
    fw, fh = f32(w), f32(h)
    ndc_x = (px + f32(0.5)) / fw * f32(2) - f32(1); ndc_y = (py + f32(0.5)) / fh * f32(2) - f32(1)
    fov = g1("tan", f32(45.0 * 1.5) * f32(0.017453292519943295) / f32(2))
    v = norm(np.stack([ndc_x * fov, ndc_y * fov / (fw / fh), np.full_like(ndc_x, -1)], -1).astype(f32))
    d = np.stack([(x[i] * v[..., 0] + y[i] * v[..., 1]) + z[i] * v[..., 2] for i in range(3)], -1).astype(f32)
    d = d.reshape(-1, 3); n_px = len(d)
    o = np.broadcast_to(eye, d.shape)
    de = lambda p: _np_mandelbulb8(O, p)
    with np.errstate(all="ignore"):
        rs = (f32(0) - o).astype(f32)
        tc = dot(d, rs); a = dot(rs, rs) - tc * tc; r2 = f32(1.15) * f32(1.15)
        insph = ~(a > r2)
        hh = np.sqrt(np.where(insph, r2 - a, 0)).astype(f32)
        tmax = tc + hh; tt = np.maximum(f32(0), tc - hh).astype(f32)
        alive = insph.copy(); hit = np.zeros(n_px, bool)
        for _ in range(128):
            if not alive.any():
                break
            idx = np.nonzero(alive)[0]
            pos = (o[idx] + d[idx] * tt[idx, None]).astype(f32)
            dist = de(pos)
            tn = (tt[idx] + dist).astype(f32)
            tt[idx] = tn
            left = tn > tmax[idx]
            close = ~left & (dist < f32(0.001))
            hit[idx[close]] = True
            alive[idx[left | close]] = False
        hi = np.nonzero(hit)[0]
        isec = (o[hi] + d[hi] * tt[hi, None]).astype(f32)
        p0 = (isec - d[hi] * f32(0.00001)).astype(f32)
        eps = f32(0.00001)
        c0 = de(p0)
        gx = c0 - de((p0 - np.array([eps, 0, 0], f32)).astype(f32))
        gy = c0 - de((p0 - np.array([0, eps, 0], f32)).astype(f32))
        gz = c0 - de((p0 - np.array([0, 0, eps], f32)).astype(f32))
        n = norm(np.stack([gx, gy, gz], -1).astype(f32))
        occ = np.zeros(len(hi), f32)
        for wgt, dl in ((f32(0.5), f32(0.016)), (f32(0.25), f32(0.081))):
            dd = de((isec + n * dl).astype(f32))
            term = f32(1) - dd / dl
            term = np.where(term < 0, f32(0), term); term = np.where(term > 1, f32(1), term)      # clamp = min(max(x,0),1)
            occ = (occ + wgt * term).astype(f32)
        occ = f32(1) - occ; occ = occ - f32(0.29); occ = occ * f32(3.5); occ = occ * occ
        occ = np.where(occ < 0, f32(0), occ); ao = np.where(occ > 1, f32(1), occ).astype(f32)
        cosi = dot(-d[hi], n); eta, kk = f32(0.4), f32(0.8)
        tmp = (eta * eta + kk * kk) * cosi * cosi
        rpar = (tmp - (f32(2) * eta * cosi) + f32(1)) / (tmp + (f32(2) * eta * cosi) + f32(1))
        tf = eta * eta + kk * kk
        rper = (tf - (f32(2) * eta * cosi) + cosi * cosi) / (tf + (f32(2) * eta * cosi) + cosi * cosi)
        fres = ((rpar + rper) / f32(2)).astype(f32)
        one = np.ones(len(hi), f32)
        col = []
        for dc, sc in ((f32(1), f32(0.8)), (f32(0.8), f32(0.8)), (f32(0.8), f32(1))):
            t1 = one * dc * f32(0.5)
            t2 = one * sc * f32(5) * fres * f32(0.5)
            t3 = one * f32(0.5) * fres * f32(0.1)
            col.append((((t1 + t2) + t3) * f32(3) * ao).astype(f32))
        color = np.ones((n_px, 3), f32)                         # misses: texture(env_reflection, dir) = 1
        color[hi] = np.stack(col, -1)
    gam = O.glsl("pow", color.reshape(-1), f32(1.0) / f32(2.2)).reshape(-1, 3)
    u8 = O.glsl("unorm8", gam.reshape(-1)).reshape(-1, 3).astype(np.uint32)
    fb = u8[:, 0] | (u8[:, 1] << 8) | (u8[:, 2] << 16) | np.uint32(0xFF000000)
    return fb.reshape(h, w), hit.reshape(h, w)


@pytest.mark.parametrize("t", [0.0, 7.0])
def test_whole_pipeline_matches_numpy_restatement(O, t):
    """The C oracle's full per-pixel pipeline against an independent numpy restatement of the shader:
    identical RGBA8 words and hit mask (bit for bit — both are IEEE-only and share the built-ins)."""
    w, h = 40, 32
    fb, hit = _np_render_mandelbulb_white(O, w, h, t)
    ref = O.render(O.SCENE_MB_POWER8, w, h, t)
    assert np.array_equal(hit, ref["hit"].astype(bool))
    assert np.array_equal(fb, ref["fb"]), int((fb != ref["fb"]).sum())
