"""GPU parity of the env-map preprocessing kernels (csrc/envmap.cu) against the oracle's
restatement of HDREnvMap.hs, through the C ABI.  resize is IEEE-only → bit-exact; the
convolution re-orders the FP32 sum (and uses squarings instead of powf) → rtol stated;
the cube build uses CUDA's acosf/atan2f instead of glibc's → ≤ 1 half-ulp on rare texels."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONV_RTOL = 1e-4      # "compare with rtol ~1e-4 pre-RGBE"


def test_resize_bit_exact(sr, O, uffizi):
    for dstw in (256, 100, 33):
        assert np.array_equal(sr.resizeHDRImage(uffizi, dstw), O.resize(uffizi, dstw)), dstw


def test_convolve_reference_powers(sr, O, uffizi):
    """BASELINE config 2: uffizi_512 → 256×128 → cos^{1,8,64,512} in one pass."""
    small = O.resize(uffizi, 256)
    got = sr.cosineConvolveHDREnvMapSweep(small, [1.0, 8.0, 64.0, 512.0])
    for i, p in enumerate((1.0, 8.0, 64.0, 512.0)):
        ref = O.convolve(small, p)
        rel = np.abs(got[i] - ref) / np.maximum(np.abs(ref), 1e-12)
        print("power %5.0f: max rel err %.2e, mean rel err %.2e" % (p, rel.max(), rel.mean()))
        assert rel.max() <= CONV_RTOL, (p, rel.max())
        # after the RGBE cache-file quantisation at most a mantissa step apart, on a tiny fraction
        qa, qb = O.rgbe_roundtrip(got[i]), O.rgbe_roundtrip(ref)
        assert (qa != qb).mean() < 0.02
        assert (np.abs(qa - qb) <= np.maximum(qb, 1e-12) * (1.0 / 64.0)).all()


def test_convolve_generic_powers_and_small_maps(sr, O, uffizi):
    tiny = O.resize(uffizi, 48)                                   # 48×24, ragged w.r.t. the 32×8 CTA tile
    got = sr.cosineConvolveHDREnvMapSweep(tiny, [2.0, 5.5, 1.0, 8.0, 64.0, 512.0, 30.0])
    for i, p in enumerate((2.0, 5.5, 1.0, 8.0, 64.0, 512.0, 30.0)):
        ref = O.convolve(tiny, p)
        assert np.allclose(got[i], ref, rtol=CONV_RTOL, atol=0), p
    assert np.allclose(sr.cosineConvolveHDREnvMap(tiny, 8.0), O.convolve(tiny, 8.0), rtol=CONV_RTOL)
    # the six-colour debug map (HDREnvMap.hs:55-73)
    dbg = O.resize(O.test_latlong_envmap(), 64)
    assert np.allclose(sr.cosineConvolveHDREnvMapSweep(dbg, [1.0, 8.0, 64.0, 512.0])[1], O.convolve(dbg, 8.0), rtol=CONV_RTOL, atol=1e-7)


def test_cube_build(sr, O, R, uffizi):
    small = O.resize(uffizi, 256)
    for slot, img in ((R.ENV_REFLECTION, uffizi), (R.ENV_COS_1, small)):
        sr.latLongHDREnvMapToCubeMap(slot, img)
        got, ref = sr.getCubeMap(slot), O.latlong_to_cube(img)
        assert got.shape == ref.shape == (6, img.shape[1] // 3, img.shape[1] // 3, 3)
        diff = got != ref
        print("cube %d: %.4f%% texels differ" % (slot, 100 * diff.mean()))
        assert diff.mean() < 0.01
        # never more than one half-precision step
        assert (np.abs(got - ref) <= np.abs(ref) * 2.0 ** -10 + 1e-7).all()
    dbg = O.test_latlong_envmap()
    sr.latLongHDREnvMapToCubeMap(R.ENV_COS_8, dbg)
    assert np.array_equal(sr.getCubeMap(R.ENV_COS_8)[:, 40:130, 40:130], O.latlong_to_cube(dbg)[:, 40:130, 40:130])


def test_bind_cube_rounds_to_half_and_rebinds(sr, R):
    rng = np.random.RandomState(5)
    faces = rng.uniform(0, 20, size=(6, 7, 7, 3)).astype(np.float32)
    sr.bindCubeMap(R.ENV_COS_64, faces)
    assert np.array_equal(sr.getCubeMap(R.ENV_COS_64), faces.astype(np.float16).astype(np.float32))
    faces2 = rng.uniform(0, 1, size=(6, 7, 7, 3)).astype(np.float32)
    sr.bindCubeMap(R.ENV_COS_64, faces2)                           # same size: storage re-used
    assert np.array_equal(sr.getCubeMap(R.ENV_COS_64), faces2.astype(np.float16).astype(np.float32))
    sr.bindCubeMap(R.ENV_COS_64, faces[:, :3, :3].copy())          # size change
    assert sr.getCubeMap(R.ENV_COS_64).shape == (6, 3, 3, 3)
    sr.unbindCubeMap(R.ENV_COS_64)
    assert sr.getCubeMap(R.ENV_COS_64) is None
    with pytest.raises(R.RmdfError):
        sr.bindCubeMap(7, faces)


def test_with_shader_renderer_builds_the_cache(R, O, tmp_path):
    """withShaderRenderer's env half (ShaderRendering.hs:66-91,131-149): cache files with the
    reference's names, skipped when present, then five cube maps bound."""
    from conftest import UFFIZI
    with R.withShaderRenderer(UFFIZI, device=0, cache_dir=tmp_path) as sr:
        names = sorted(os.listdir(tmp_path))
        assert names == ["uffizi_512_cache_pow_%s.hdr" % p for p in ("1.0", "512.0", "64.0", "8.0")]
        assert sr.srEnvCubeMaps == ["env_reflection", "env_cos_1", "env_cos_8", "env_cos_64", "env_cos_512"]
        assert sr.getCubeMap(R.ENV_REFLECTION).shape[1] == 170 and sr.getCubeMap(R.ENV_COS_512).shape[1] == 85
        cached = R.loadHDRImage(tmp_path / "uffizi_512_cache_pow_8.0.hdr")
        ref = O.rgbe_roundtrip(O.convolve(O.resize(O.hdr_decode(UFFIZI.read_bytes()), 256), 8.0))
        assert (cached != ref).mean() < 0.02
        mt = os.path.getmtime(tmp_path / "uffizi_512_cache_pow_8.0.hdr")
    with R.withShaderRenderer(UFFIZI, device=0, cache_dir=tmp_path):
        assert os.path.getmtime(tmp_path / "uffizi_512_cache_pow_8.0.hdr") == mt       # cache hit: not rewritten


def test_convolve_errors(sr, R):
    with pytest.raises(R.RmdfError):
        sr.cosineConvolveHDREnvMapSweep(np.zeros((1, 1, 3), np.float32), [1.0])
    with pytest.raises(R.RmdfError):
        sr.cosineConvolveHDREnvMapSweep(np.zeros((8, 16, 3), np.float32), [])
