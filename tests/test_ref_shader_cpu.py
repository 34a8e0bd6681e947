"""oracle/_ref — the reference's OWN fragment shader (/root/reference/fragment.shd) compiled by
g++ through oracle/ref/glsl_compat.hpp — against the hand-written C oracle, bit for bit:
every function on the hot path over lattices of inputs, and whole frames per shader variant
(the four FragmentShader values of ShaderRendering.hs:119-123 plus the RAY_TRACING_TEST and
CAM_ORTHO switches).  This is what pins rows a1-a11 of SURVEY.md §8a to the reference itself.

Needs oracle/_ref/libref_fragment.so: built here from /root/reference (make -C oracle ref), or
travelling prebuilt to the GPU box.  Where neither exists the tests skip and
tests/golden/ref_*.npz (generated from the same library, tests/golden/make_ref_golden.py) still
pin the oracle through test_ref_golden_*."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
f32 = np.float32


@pytest.fixture(scope="module")
def REF(O):
    from oracle import oracle_ref
    if not oracle_ref.available():
        pytest.skip("oracle/_ref/libref_fragment.so absent and /root/reference not here to build it")
    return oracle_ref


def bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


def same(a, b):
    """bit-identical; a NaN must meet a NaN (sign and payload of a NaN carry no meaning: the k3 = 0
    axis of triplex_pow8, fragment.shd:92, is inf·0 on both sides)"""
    a, b = np.ascontiguousarray(a, f32), np.ascontiguousarray(b, f32)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(bits(a)[~na], bits(b)[~nb])


def lattice(n=9, lim=1.2, seed=0):
    g = np.linspace(-lim, lim, n, dtype=f32)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.RandomState(seed)
    return np.concatenate([pts, rng.uniform(-lim, lim, size=(4000, 3)).astype(f32),
                           rng.normal(0, 0.4, size=(2000, 3)).astype(f32)])


SCENES = [0, 1, 2, 3]


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("t", [0.0, 7.0, 11.3])
def test_distance_estimator_lattice(O, REF, scene, t):
    pts = lattice(seed=scene)
    assert same(REF.distance(scene, t, pts), O.distance(scene, t, pts))


def test_triplex_powers_and_primitives(O, REF):
    pts = lattice(seed=5)
    assert same(REF.triplex_pow8(pts), O.triplex_pow(pts, 8.0, True))
    for p in (2.0, 3.7, 6.5):
        assert same(REF.triplex_pow(pts, p), O.triplex_pow(pts, p, False))
    rng = np.random.RandomState(2)
    a, b = rng.uniform(-0.5, 1.5, 5000).astype(f32), rng.uniform(-0.5, 1.5, 5000).astype(f32)
    assert same(REF.smin(a, b, 64.0), O.smin(a, b, 64.0))
    assert same(REF.smin(a, b, 32.0), O.smin(a, b, 32.0))
    cv = O.cornell_vertices().reshape(32, 9)
    for tri in cv[::5]:
        assert same(REF.de_triangle(pts, tri), O.de_triangle(pts, tri))
    c = np.linspace(-1, 1, 4001, dtype=f32)
    assert same(REF.fresnel_conductor(c, 0.4, 0.8), O.fresnel_conductor(c, 0.4, 0.8))


def _rays(scene, t, O, n=24):
    cam = O.camera(scene, t)
    rays = np.array([O.generate_ray(cam, n, n, x, y) for y in range(n) for x in range(n)], f32)
    return cam, rays[:, :3], rays[:, 3:]


@pytest.mark.parametrize("scene", SCENES)
def test_camera_and_rays(O, REF, scene):
    for t in (0.0, 7.0):
        cam = O.camera(scene, t)
        # main() builds the eye (fragment.shd:884-898) and calls lookat(eye, 0, +Y); whole frames
        # below cover the eye, here lookat and generate_ray on their own
        assert same(REF.lookat(cam[9:12], variant=scene), cam)
        for (w, h) in ((24, 24), (37, 19)):
            for (px, py) in ((0, 0), (w - 1, h - 1), (w // 2, h // 3), (5, 11)):
                assert same(REF.generate_ray(cam, w, h, px, py, False, 67.5, variant=scene), O.generate_ray(cam, w, h, px, py, False))
                assert same(REF.generate_ray(cam, w, h, px, py, True, 2.0, variant=scene), O.generate_ray(cam, w, h, px, py, True))


@pytest.mark.parametrize("scene", SCENES)
def test_ray_sphere_march_normal_ao(O, REF, scene):
    radius = {2: 1.15, 3: 1.5}.get(scene, 1.0)
    for t in (0.0, 7.0):
        _, org, dirs = _rays(scene, t, O)
        assert same(REF.ray_sphere(org, dirs, radius, variant=scene), O.ray_sphere(org, dirs, radius))
        rm_ref, rm = REF.ray_march(scene, t, org, dirs), O.ray_march(scene, t, org, dirs)
        assert same(rm_ref, rm)
        hit = rm[:, 0] > 0
        assert hit.sum() > 20
        isec = (org[hit] + dirs[hit] * rm[hit, 1:2]).astype(f32)
        for central in (False, True):
            assert same(REF.normal(scene, t, isec, central), O.normal(scene, t, isec, central))
        n = O.normal(scene, t, isec)
        n = np.where(np.isfinite(n), n, f32(0.0))
        assert same(REF.distance_ao(scene, t, isec, n), O.distance_ao(scene, t, isec, n))


def _synthetic_envs(O):
    rng = np.random.RandomState(1234)
    envs = []
    for w in (6, 4, 4, 3, 3):
        e = rng.uniform(0.0, 2.0, size=(6, w, w, 3)).astype(f32)
        envs.append(O.glsl("round_to_half", e).reshape(e.shape))
    return envs


def _frames_equal(a, b):
    return np.array_equal(a["fb"], b["fb"]) and same(a["linear"], b["linear"]) and np.array_equal(a["hit"], b["hit"])


@pytest.mark.parametrize("scene", SCENES)
def test_frames_white_env(O, REF, scene):
    for t in (0.0, 7.0):
        assert _frames_equal(REF.render(scene, 64, 64, t), O.render(scene, 64, 64, t))
    assert _frames_equal(REF.render(scene, 45, 31, 3.0), O.render(scene, 45, 31, 3.0))          # odd sizes: partial quads


@pytest.mark.parametrize("scene", SCENES)
def test_frames_env_lit_both_lod_policies(O, REF, scene, oracle_envs):
    syn = _synthetic_envs(O)
    for envs, (w, h) in ((syn, (64, 64)), (oracle_envs, (96, 64))):
        for pol in (O.LOD_QUAD, O.LOD_ALWAYS_LINEAR):
            a, b = REF.render(scene, w, h, 7.0, envs=envs, lod_policy=pol), O.render(scene, w, h, 7.0, envs=envs, lod_policy=pol)
            assert _frames_equal(a, b)
            assert a["nearest_quads"] == b["counters"]["nearest_quads"]


def test_frames_shader_switches(O, REF, oracle_envs):
    for scene, flags in ((0, O.FLAG_RAY_TRACING_TEST), (1, O.FLAG_CAM_ORTHO), (2, O.FLAG_CAM_ORTHO)):
        for t in (0.0, 1.0):
            assert _frames_equal(REF.render(scene, 80, 60, t, envs=oracle_envs, flags=flags), O.render(scene, 80, 60, t, envs=oracle_envs, flags=flags))


def test_crops_of_the_baseline_frames(O, REF, oracle_envs):
    """crops of BASELINE configs 1, 3 and 4 at their real frame-buffer sizes"""
    cases = [(2, 512, 512, None, [(200, 200, 264, 232)]),
             (0, 3840, 2160, oracle_envs, [(1900, 1000, 1932, 1016), (3808, 2144, 3840, 2160)]),
             (1, 3840, 2160, oracle_envs, [(1700, 1200, 1764, 1232)]),
             (2, 15360, 8640, oracle_envs, [(7000, 4000, 7064, 4032), (0, 0, 32, 32), (15328, 8608, 15360, 8640), (5001, 3001, 5033, 3017)])]
    for scene, w, h, envs, rects in cases:
        for r in rects:
            assert _frames_equal(REF.render(scene, w, h, 0.0, envs=envs, rect=r), O.render(scene, w, h, 0.0, envs=envs, rect=r))


def test_reference_reproduces_the_committed_golden_frames(O, REF):
    """tests/golden/frames_64.npz was minted from the oracle in round 1; the reference's shader gives the same words"""
    z = np.load(GOLD / "frames_64.npz")
    envs = _synthetic_envs(O)
    for s in range(4):
        for t in (0.0, 7.0):
            assert np.array_equal(REF.render(s, 64, 64, t)["fb"], z["white_s%d_t%d" % (s, int(t))])
            assert np.array_equal(REF.render(s, 64, 64, t, envs=envs)["fb"], z["env_s%d_t%d" % (s, int(t))])
