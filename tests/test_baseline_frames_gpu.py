"""-m gpu: BASELINE.json's configurations at their REAL frame sizes.  The whole frame is rendered
once on the GPU through the C ABI; crops spread over it (corners, centre, last row and column, odd
offsets, the region where px + py·w approaches 2^31) are then compared with the CPU oracle's
render of the same rectangle (oracle.render(rect=…) / oracle.julia(rows=…) + the glGenerateMipmap
resolve).  Parity build: bit-identical.  Fast build: BASELINE's tolerance, measured and printed.

Also: the CUDA kernels against outputs of the reference's OWN fragment shader — the committed
fixtures tests/golden/ref_shader.npz and, when oracle/_ref/libref_fragment.so travelled to this
box, the compiled shader itself."""
import numpy as np
import pytest

from conftest import psnr

pytestmark = pytest.mark.gpu


def _bind(sr, envs):
    for s in range(5):
        if envs is None:
            sr.unbindCubeMap(s)
        else:
            sr.bindCubeMap(s, envs[s])


def _crops(out_w, out_h, cw, ch, n_random=10, seed=0):
    """(x0, y0) of output-pixel crops: four corners, centre, last row/column strips, random ones"""
    rng = np.random.RandomState(seed)
    pts = [(0, 0), (out_w - cw, 0), (0, out_h - ch), (out_w - cw, out_h - ch), ((out_w - cw) // 2, (out_h - ch) // 2),
           ((out_w - cw) // 2 + 1, out_h - ch), (out_w - cw, (out_h - ch) // 2 + 1)]
    for _ in range(n_random):
        pts.append((int(rng.randint(0, out_w - cw)), int(rng.randint(0, out_h - ch))))
    return pts


def _check_frame_crops(sr, O, R, scene, fb_w, fb_h, k, envs, cw, ch, time=0.0, n_random=10, only_geometry=False):
    s = 1 << k
    out_w, out_h = fb_w >> k, fb_h >> k
    _bind(sr, envs)
    got = sr.drawShaderTile(R.FragmentShader(scene), None, fb_w, fb_h, time, resolve_log2=k)
    fast = sr.drawShaderTile(R.FragmentShader(scene), None, fb_w, fb_h, time, resolve_log2=k, mode=R.MODE_FAST)
    assert got.shape == (out_h, out_w)
    crops = _crops(out_w, out_h, cw, ch, n_random, seed=scene)
    if only_geometry:       # keep the crops where something is on screen plus the mandatory corners
        crops = crops[:7] + [((out_w - cw) // 2 + dx, (out_h - ch) // 2 + dy) for dx in (-out_w // 8, 0, out_w // 8) for dy in (-out_h // 6, 0, out_h // 6)]
    bad, se, npx = 0, 0.0, 0
    for (x0, y0) in crops:
        rect = (x0 * s, y0 * s, (x0 + cw) * s, (y0 + ch) * s)
        ref = O.render(scene, fb_w, fb_h, time, envs=envs, rect=rect, want_linear=False)
        want = O.mip_resolve(ref["fb"], k) if k else ref["fb"]
        g = got[y0:y0 + ch, x0:x0 + cw]
        assert np.array_equal(g, want), "parity build differs from the oracle in crop (%d, %d): %d pixels" % (x0, y0, int((g != want).sum()))
        f8 = fast[y0:y0 + ch, x0:x0 + cw].view(np.uint8).reshape(ch, cw, 4)[..., :3].astype(np.float64)
        w8 = want.view(np.uint8).reshape(ch, cw, 4)[..., :3].astype(np.float64)
        se += float(((f8 - w8) ** 2).sum()); npx += f8.size
    p = float("inf") if se == 0 else 10 * np.log10(255.0 ** 2 * npx / se)
    print("scene %d %dx%d k=%d: %d crops bit-exact (parity); fast build gamma-space PSNR over the crops %.2f dB" % (scene, fb_w, fb_h, k, len(crops), p))
    return p


def test_config1_full_frame(sr, O, R):
    """configs[0]: Mandelbulb power 8, 512×512, 1 spp, no env lighting — the whole frame."""
    _bind(sr, None)
    ref = O.render(2, 512, 512, 0.0)
    got = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, 512, 512, 0.0)
    assert np.array_equal(got, ref["fb"])


@pytest.mark.parametrize("scene", [0, 1])
def test_config3_crops_3840x2160(sr, O, R, oracle_envs, scene):
    """configs[2]: Cornell box / CSG scene, 1920×1080 × 4 spp (frame buffer 3840×2160, k = 1)."""
    p = _check_frame_crops(sr, O, R, scene, 3840, 2160, 1, oracle_envs, 32, 16)
    assert p >= 30.0


def test_config4_crops_15360x8640(sr, O, R, oracle_envs):
    """configs[3]: Mandelbulb, 3840×2160 × 16 spp (frame buffer 15360×8640, k = 2), env-lit."""
    p = _check_frame_crops(sr, O, R, 2, 15360, 8640, 2, oracle_envs, 32, 16, only_geometry=True)
    assert p >= 35.0


def test_config4_fast_build_tolerance_on_oracle_crops(sr, O, R, oracle_envs):
    """BASELINE's bar for the fast build at config 4's sampling rate, AGAINST THE ORACLE (not the parity
    build): hit masks agree on >= 99.9 % of samples and linear-RGB PSNR >= 50 dB on the 16-spp resolved
    image.  The bar is stated on the frame, so the estimate uses a UNIFORM sample of it: a 16×8 grid of
    128×64-sample crops (two per tile of the reference's 8×8 grid), level-0 FP32 colour of the real
    15360×8640 frame, 4×4 box in linear RGB clamped to [0,1].  The figure over the crops that contain
    geometry only is printed beside it (harsher: no background pixels to dilute the error)."""
    _bind(sr, oracle_envs)
    fb_w, fb_h, s = 15360, 8640, 4
    tw, th = fb_w // 8, fb_h // 8
    cw, ch = 128, 64
    box = lambda a: np.clip(np.nan_to_num(a), 0, 1).reshape(ch // s, s, cw // s, s, 3).mean(axis=(1, 3))
    se = n = se_g = n_g = 0.0
    agree, cnt = 0.0, 0
    for idx in range(64):
        fast = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, idx, fb_w, fb_h, 0.0, out_format=R.OUT_RGBA32F,
                                 mode=R.MODE_FAST, layout=R.LAYOUT_PACKED_TILES).reshape(th, tw, 4)
        tx, ty = idx % 8, idx // 8
        for cx in (tw // 4 - cw // 2, 3 * tw // 4 - cw // 2):
            cy = th // 2 - ch // 2
            rect = (tx * tw + cx, ty * th + cy, tx * tw + cx + cw, ty * th + cy + ch)
            ref = O.render(2, fb_w, fb_h, 0.0, envs=oracle_envs, rect=rect)
            f = fast[cy:cy + ch, cx:cx + cw]
            d = box(f[..., :3]).astype(np.float64) - box(ref["linear"]).astype(np.float64)
            se += float((d ** 2).sum()); n += d.size
            if ref["hit"].any():
                se_g += float((d ** 2).sum()); n_g += d.size
            agree += float(np.mean((f[..., 3] > 0.5) == ref["hit"].astype(bool))); cnt += 1
    p = float("inf") if se == 0 else 10 * np.log10(n / se)
    pg = float("inf") if se_g == 0 else 10 * np.log10(n_g / se_g)
    print("config 4 fast build vs oracle, 128 uniform crops of 128x64 samples: hit agreement %.5f%%, linear PSNR %.2f dB "
          "(crops with geometry only: %.2f dB over %d of %d output pixels)" % (100 * agree / cnt, p, pg, n_g // 3, n // 3))
    assert agree / cnt >= 0.999 and p >= 50.0


@pytest.mark.parametrize("smooth", [False, True])
def test_config5_julia_crops_61440x34560(sr, O, smooth):
    """configs[4]: Julia set 7680×4320 × 64 spp (frame buffer 61440×34560 = 2.12e9 samples, k = 3).
    Row bands of the resolved frame vs oracle.julia(rows=…) + mip resolve, including the last rows,
    where px + py·w reaches 2.12e9 (1 % under 2^31)."""
    fb_w, fb_h, k, tick = 61440, 34560, 3, 7.0
    got = sr.juliaAnimated(fb_w, fb_h, None, smooth, tick, resolve_log2=k)
    assert got.shape == (4320, 7680)
    rng = np.random.RandomState(5)
    rows = [0, 1, 4319, 4318, 2160, 2159, 1080, 3239, 4200] + [int(r) for r in rng.randint(0, 4320, 8)]
    diff = 0
    for oy in rows:
        ref = O.julia(fb_w, fb_h, smooth, tick, rows=(oy * 8, oy * 8 + 8))
        want = O.mip_resolve(ref, k)
        d = int((got[oy:oy + 1] != want).sum())
        if not smooth:
            assert d == 0, "row %d: %d pixels differ" % (oy, d)
        diff += d
    print("julia smooth=%s: %d row bands of 61440x8 samples checked, %d resolved pixels differ" % (smooth, len(rows), diff))
    assert diff <= 0.001 * len(rows) * 7680            # smooth: CUDA vs glibc logf; measured 0


def test_config2_prefilter_sweep(sr, O, R, uffizi):
    """configs[1]: uffizi_512.hdr → 256×128 → cos^{1,8,64,512}: the real sweep against the sequential FP32 oracle."""
    small = sr.resizeHDRImage(uffizi, 256)
    assert np.array_equal(small, O.resize(uffizi, 256))
    conv = sr.cosineConvolveHDREnvMapSweep(small, sr.powers)
    for i, p in enumerate(sr.powers):
        want = O.convolve(small, p)
        np.testing.assert_allclose(conv[i], want, rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------- the reference's own shader
def _synthetic_envs(O):
    rng = np.random.RandomState(1234)
    envs = []
    for w in (6, 4, 4, 3, 3):
        e = rng.uniform(0.0, 2.0, size=(6, w, w, 3)).astype(np.float32)
        envs.append(O.glsl("round_to_half", e).reshape(e.shape))
    return envs


def test_cuda_equals_reference_shader_golden(sr, O, R):
    """CUDA parity build == tests/golden/ref_shader.npz, i.e. outputs of /root/reference/fragment.shd
    compiled by g++ in the dev container (tests/golden/make_ref_golden.py)."""
    from pathlib import Path
    z = np.load(Path(__file__).resolve().parent / "golden" / "ref_shader.npz")
    envs = _synthetic_envs(O)
    for s in range(4):
        for t in (0.0, 7.0):
            for name, e in (("white", None), ("env", envs)):
                _bind(sr, e)
                k = "%s_s%d_t%d" % (name, s, int(t))
                assert np.array_equal(sr.drawShaderTile(R.FragmentShader(s), None, 64, 48, t), z["fb_" + k]), k
                lin = sr.drawShaderTile(R.FragmentShader(s), None, 64, 48, t, out_format=R.OUT_RGBA32F)
                want = z["lin_" + k].view(np.float32)
                nan = np.isnan(want)
                assert np.array_equal(np.isnan(lin[..., :3]), nan) and np.array_equal(lin[..., :3][~nan].view(np.uint32), want[~nan].view(np.uint32)), k
                assert np.array_equal(np.packbits((lin[..., 3] > 0.5).astype(np.uint8)), z["hit_" + k]), k
    _bind(sr, envs)
    for s, fl in ((0, 4), (1, 2), (2, 2)):
        assert np.array_equal(sr.drawShaderTile(R.FragmentShader(s), None, 64, 48, 1.0, flags=fl), z["fb_flag%d_s%d" % (fl, s)])
    _bind(sr, None)
    for key, scene, w, h, rect in (("fb_c1_crop", 2, 512, 512, (224, 240, 288, 272)), ("fb_c3_cornell_crop", 0, 3840, 2160, (1888, 1064, 1952, 1096)),
                                   ("fb_c3_csg_crop", 1, 3840, 2160, (1888, 1064, 1952, 1096)), ("fb_c4_crop", 2, 15360, 8640, (7648, 4304, 7712, 4336))):
        full = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, 0.0)
        assert np.array_equal(full[rect[1]:rect[3], rect[0]:rect[2]], z[key]), key


def test_cuda_equals_reference_shader_live(sr, O, R, oracle_envs):
    """CUDA parity build == the compiled reference shader (oracle/_ref), env-lit, all four variants."""
    from oracle import oracle_ref as REF
    if not REF.available():
        pytest.skip("oracle/_ref/libref_fragment.so did not travel to this box")
    _bind(sr, oracle_envs)
    for scene in range(4):
        for (w, h, t) in ((256, 160, 0.0), (97, 61, 7.0)):
            ref = REF.render(scene, w, h, t, envs=oracle_envs)
            assert np.array_equal(sr.drawShaderTile(R.FragmentShader(scene), None, w, h, t), ref["fb"]), (scene, w, h)
            lin = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, t, out_format=R.OUT_RGBA32F)
            assert np.array_equal(lin[..., 3] > 0.5, ref["hit"].astype(bool))
    # a crop of the headline frame
    rect = (7744, 3400, 8000, 3528)                    # inside tile 28 = (tx 4, ty 3): x in [7680, 9600), y in [3240, 4320)
    ref = REF.render(2, 15360, 8640, 0.0, envs=oracle_envs, rect=rect)
    got = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, 28, 15360, 8640, 0.0, layout=R.LAYOUT_PACKED_TILES).reshape(1080, 1920)
    x0, y0 = rect[0] - 4 * 1920, rect[1] - 3 * 1080
    assert np.array_equal(got[y0:y0 + 128, x0:x0 + 256], ref["fb"])
