"""Accuracy of include/rmdf_glsl.h (the deterministic GLSL built-ins shared by the parity
kernel and the oracle) against libm in double precision.  CPU only."""
import numpy as np

f32 = np.float32


def _ulps(got, ref64):
    ref32 = ref64.astype(f32)
    ulp = np.abs(np.nextafter(ref32, f32(np.inf)).astype(np.float64) - ref32.astype(np.float64))
    ulp = np.maximum(ulp, 1e-45)
    return np.abs(got.astype(np.float64) - ref64) / ulp


def test_log_exp_pow(O):
    rng = np.random.RandomState(0)
    x = np.exp(rng.uniform(-20, 20, 400000)).astype(f32)
    assert _ulps(O.glsl("log", x), np.log(x.astype(np.float64))).max() <= 1.0
    x = rng.uniform(-100, 88, 400000).astype(f32)
    assert _ulps(O.glsl("exp", x), np.exp(x.astype(np.float64))).max() <= 1.0
    x = rng.uniform(0, 4, 400000).astype(f32)
    g = f32(1.0) / f32(2.2)
    assert _ulps(O.glsl("pow", x, g), np.power(x.astype(np.float64), np.float64(g))).max() <= 8.0
    # special values
    assert O.glsl("log", [0.0])[0] == -np.inf and np.isnan(O.glsl("log", [-1.0])[0])
    assert O.glsl("log", [1e-40])[0] == f32(np.log(np.float64(f32(1e-40))))
    assert O.glsl("exp", [-200.0])[0] == 0 and O.glsl("exp", [89.0])[0] == np.inf
    assert O.glsl("exp", [-103.0])[0] > 0                       # denormal result, not flushed
    assert O.glsl("pow", [0.0, -1.0], 0.4545)[0] == 0 and O.glsl("pow", [0.0, -1.0], 0.4545)[1] == 0
    assert np.isnan(O.glsl("pow", [np.nan], 0.4545)[0])


def test_trig(O):
    rng = np.random.RandomState(1)
    x = rng.uniform(-30, 30, 400000).astype(f32)
    for fn, ref in (("sin", np.sin), ("cos", np.cos)):
        r = ref(x.astype(np.float64))
        err = np.abs(O.glsl(fn, x).astype(np.float64) - r)
        assert (err <= 2.0 * np.maximum(np.abs(r) * 2.0 ** -23, 2.0 ** -24)).all()
    x = rng.uniform(-8000, 8000, 200000).astype(f32)
    assert np.abs(O.glsl("sin", x) - np.sin(x.astype(np.float64))).max() < 3e-7
    y, x = rng.uniform(-2, 2, 400000).astype(f32), rng.uniform(-2, 2, 400000).astype(f32)
    assert _ulps(O.glsl("atan2", y, x), np.arctan2(y.astype(np.float64), x.astype(np.float64))).max() <= 4.0
    x = rng.uniform(-1, 1, 400000).astype(f32)
    assert _ulps(O.glsl("acos", x), np.arccos(x.astype(np.float64))).max() <= 4.0
    assert O.glsl("atan2", [0.0], [0.0])[0] == 0
    assert O.glsl("acos", [1.0])[0] == 0 and abs(O.glsl("acos", [-1.0])[0] - np.pi) < 1e-6


def test_ieee_pieces(O):
    rng = np.random.RandomState(2)
    x = np.exp(rng.uniform(-30, 30, 100000)).astype(f32)
    assert np.array_equal(O.glsl("inversesqrt", x), (f32(1) / np.sqrt(x)).astype(f32))     # sqrt.rn then div.rn
    x = np.concatenate([rng.uniform(-70000, 70000, 100000), rng.uniform(-1e-4, 1e-4, 100000), [65504, 65519.9, 65520, 1e-8, 0, -0.0]]).astype(f32)
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).astype(f32)                                             # numpy: RTNE like the GL upload
    assert np.array_equal(O.glsl("round_to_half", x).view(np.uint32), want.view(np.uint32))
    c = np.array([-1, 0, 0.5 / 255, 0.4999 / 255, 1.0, 2.0, np.nan, 0.5], f32)
    assert O.glsl("unorm8", c).tolist() == [0, 0, 1, 0, 255, 255, 0, 128]
    assert np.allclose(O.glsl("mod", [7.0, -1.0, 3.5 / 2], 9.0), [7.0, 8.0, 1.75])
