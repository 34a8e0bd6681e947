"""GPU parity of the march kernel (csrc/march_kernel.cuh) against the CPU oracle, through
the C ABI.  Parity mode must be BIT-IDENTICAL (RGBA8 words, linear FP32 colour, hit mask,
work counters); fast mode is held to the tolerance BASELINE.json states: hit/miss masks
agree on >= 99.9 % of pixels, and its PSNR is reported next to the intrinsic noise floor."""
import numpy as np
import pytest

from conftest import psnr

pytestmark = pytest.mark.gpu

SCENES = [0, 1, 2, 3]


def _bind(sr, envs):
    for s in range(5):
        if envs is None:
            sr.unbindCubeMap(s)
        else:
            sr.bindCubeMap(s, envs[s])


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("time", [0.0, 7.0])
def test_rgba8_bit_exact_white_env(sr, O, R, scene, time):
    """BASELINE config 1 shape (no env lighting = all five maps constant 1.0), small frame."""
    w, h = 160, 96
    _bind(sr, None)
    ref = O.render(scene, w, h, time)
    got = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, time)
    assert np.array_equal(got, ref["fb"]), "mismatching pixels: %d" % int((got != ref["fb"]).sum())


@pytest.mark.parametrize("scene", SCENES)
def test_linear_and_hitmask_bit_exact_env_lit(sr, O, R, oracle_envs, scene):
    w, h = 192, 128
    _bind(sr, oracle_envs)
    ref = O.render(scene, w, h, 7.0, envs=oracle_envs)
    sr.resetCounters()
    got = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, 7.0, out_format=R.OUT_RGBA32F)
    hit = got[..., 3] > 0.5
    assert np.array_equal(hit, ref["hit"].astype(bool))
    # NaN-aware bit comparison of the linear colour
    assert np.array_equal(got[..., :3].view(np.uint32), ref["linear"].view(np.uint32))
    c = sr.counters()
    rc = ref["counters"]
    assert (c["samples"], c["in_sphere"], c["hits"], c["march_de"], c["shade_de"]) == \
           (rc["samples"], rc["in_sphere"], rc["hits"], rc["march_de"], rc["shade_de"])
    if scene in (2, 3):
        assert c["inner_iters"] == rc["mb_iters"]
    fb = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, 7.0)
    assert np.array_equal(fb, ref["fb"])


def test_config1_mandelbulb_512(sr, O, R):
    """BASELINE config 1: Mandelbulb power 8, 512×512, 1 spp, no env lighting."""
    _bind(sr, None)
    ref = O.render(2, 512, 512, 0.0)
    got = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, 512, 512, 0.0, out_format=R.OUT_RGBA32F)
    assert np.array_equal(got[..., 3] > 0.5, ref["hit"].astype(bool))
    assert psnr(np.clip(got[..., :3], 0, 1), np.clip(ref["linear"], 0, 1)) == float("inf")
    fb = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, 512, 512, 0.0)
    assert np.array_equal(fb, ref["fb"])


def test_always_linear_policy(sr, O, R, oracle_envs):
    _bind(sr, oracle_envs)
    ref = O.render(2, 128, 128, 0.0, envs=oracle_envs, lod_policy=O.LOD_ALWAYS_LINEAR)
    got = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, 128, 128, 0.0, lod_policy=R.LOD_ALWAYS_LINEAR)
    assert np.array_equal(got, ref["fb"])


def test_odd_frame_size(sr, O, R, oracle_envs):
    """Ragged input: odd dimensions (partial 2×2 quads, partial 8×4 tiles, partial blocks)."""
    _bind(sr, oracle_envs)
    for (w, h) in [(67, 45), (1, 1), (33, 5)]:
        ref = O.render(1, w, h, 3.0, envs=oracle_envs)
        got = sr.drawShaderTile(R.FragmentShader.FSDETestShader, None, w, h, 3.0)
        assert np.array_equal(got, ref["fb"]), (w, h)


def test_tiles_compose_to_full_frame(sr, O, R, oracle_envs):
    """drawShaderTile tile by tile (ShaderRendering.hs:183-196) == the full-frame draw; a single
    tile leaves the other pixels untouched."""
    _bind(sr, oracle_envs)
    w, h = 256, 128
    full = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0)
    fb = np.full((h, w), 0xDEADBEEF, np.uint32)
    sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, 9, w, h, 0.0, fb=fb)
    tw, th = w // 8, h // 8
    assert np.array_equal(fb[th:2 * th, tw:2 * tw], full[th:2 * th, tw:2 * tw])
    mask = np.ones((h, w), bool); mask[th:2 * th, tw:2 * tw] = False
    assert (fb[mask] == 0xDEADBEEF).all()
    for idx in range(64, 128):          # idx mod 64
        sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, idx, w, h, 0.0, fb=fb)
    assert np.array_equal(fb, full)


@pytest.mark.parametrize("size", [(100, 100), (67, 45), (5, 3), (250, 131)])
def test_tiles_of_any_frame_size(sr, O, R, oracle_envs, size):
    """ShaderRendering.hs:183-195 cuts ANY frame into 8×8 NDC rectangles; a pixel belongs to the tile that contains
    its centre.  Tiles with odd edges split 2×2 fragment quads: GL shades such a quad whole (helper invocations), so
    the min/mag decision — and every pixel — equals the full-frame draw.  Tile by tile against oracle.render(rect=…),
    frame layout and packed layout, RGBA8 and linear output; the 64 tiles partition the frame."""
    _bind(sr, oracle_envs)
    w, h = size
    edge = lambda t, dim: (2 * t * dim + 7) >> 4
    full = O.render(1, w, h, 3.0, envs=oracle_envs)
    acc = np.full((h, w), 0xDEADBEEF, np.uint32)
    covered = np.zeros((h, w), np.int32)
    for idx in range(64):
        tx, ty = idx % 8, idx // 8
        x0, x1, y0, y1 = edge(tx, w), edge(tx + 1, w), edge(ty, h), edge(ty + 1, h)
        sr.drawShaderTile(R.FragmentShader.FSDETestShader, idx, w, h, 3.0, fb=acc)
        covered[y0:y1, x0:x1] += 1
        if x1 > x0 and y1 > y0:
            ref = O.render(1, w, h, 3.0, envs=oracle_envs, rect=(x0, y0, x1, y1))
            assert np.array_equal(ref["fb"], full["fb"][y0:y1, x0:x1])                 # the oracle itself: tiling does not change a pixel
            assert np.array_equal(acc[y0:y1, x0:x1], ref["fb"]), (idx, x0, x1, y0, y1)
            packed = sr.drawShaderTile(R.FragmentShader.FSDETestShader, idx, w, h, 3.0, layout=R.LAYOUT_PACKED_TILES)
            assert packed.size == (x1 - x0) * (y1 - y0) and np.array_equal(packed.reshape(y1 - y0, x1 - x0), ref["fb"])
            if idx in (0, 9, 27, 63):
                lin = sr.drawShaderTile(R.FragmentShader.FSDETestShader, idx, w, h, 3.0, layout=R.LAYOUT_PACKED_TILES, out_format=R.OUT_RGBA32F)
                assert np.array_equal(lin.reshape(y1 - y0, x1 - x0, 4)[..., :3].view(np.uint32), ref["linear"].view(np.uint32))
    assert (covered == 1).all()
    assert np.array_equal(acc, full["fb"])


@pytest.mark.parametrize("k", [1, 2, 3])
def test_in_kernel_resolve_equals_mip_chain(sr, O, R, oracle_envs, k):
    """spp = 4^k: the kernel's in-register/shared resolve == glGenerateMipmap level k of level 0."""
    _bind(sr, oracle_envs)
    w, h = 256, 192
    ref = O.render(2, w, h, 0.0, envs=oracle_envs)
    want = O.mip_resolve(ref["fb"], k)
    got = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
    assert got.shape == (h >> k, w >> k)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("count", [2, 4, 8])
@pytest.mark.parametrize("grid", [None, ((16, 8), 2), ((32, 16), 2), ((8, 32), 3)])
def test_shards_partition_the_frame(sr, R, oracle_envs, count, grid):
    """Multi-GPU decomposition on one GPU: the packed tiles of all ranks reassemble the frame,
    for the reference's 8×8 grid and for finer skewed shard grids."""
    from importlib import import_module
    par = import_module("ray-marching-distance-fields_b200.parallel")
    _bind(sr, oracle_envs)
    w, h, k = 512, 256, 1
    g, skew = grid if grid else (par.DEFAULT_GRID, par.DEFAULT_SKEW)
    kw = dict(shard_grid=g, shard_skew=skew) if grid else {}
    full = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
    packed = [sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k,
                                shard_rank=r, shard_count=count, layout=R.LAYOUT_PACKED_TILES, **kw) for r in range(count)]
    frame = par.assemble_tiles_numpy(packed, w >> k, h >> k, count, g, skew)
    assert np.array_equal(frame, full)
    # frame layout: each rank fills only its own tiles of a shared-size buffer
    acc = np.zeros_like(full)
    for r in range(count):
        acc |= sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k,
                                 shard_rank=r, shard_count=count, **kw)
    assert np.array_equal(acc, full)


def test_peer_frame_buffer_two_gpus(R, oracle_envs):
    """The fused gather: a second GPU's kernel stores its tiles straight into GPU 0's frame buffer
    (peer memory); together with GPU 0's own tiles the frame is bit-identical to a single-GPU render.
    Needs two devices (run with gpurun --gpus 2); skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w, h, k = 1024, 512, 2
    with R.ShaderRenderer(0) as a, R.ShaderRenderer(1) as b:
        for r in (a, b):
            for s in range(5):
                r.bindCubeMap(s, oracle_envs[s])
        want = a.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
        b.enablePeer(0)
        ptr, _handle = a.sharedAlloc((w >> k) * (h >> k) * 4)
        for rank, r in enumerate((a, b)):
            p = r.makeParams(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k, shard_rank=rank, shard_count=2,
                             shard_grid=(16, 8), shard_skew=1)
            r.renderDevice(p, ptr)
        a.sync(); b.sync()
        class _W:
            pass
        wv = _W()
        wv.__cuda_array_interface__ = {"shape": ((h >> k) * (w >> k),), "typestr": "<i4", "data": (ptr, False), "version": 3}
        got = torch.as_tensor(wv, device="cuda:0").cpu().numpy().view(np.uint32).reshape(h >> k, w >> k)
        a.sharedFree(ptr)
    assert np.array_equal(got, want)


def test_row_ranges_compose_and_striped_host_path(sr, O, R, oracle_envs):
    """rows=(y0, y1): horizontal stripes of the frame (and of a rank's shard) land in place, so a
    host can pipeline a frame (copy stripe i while stripe i+1 is marched)."""
    _bind(sr, oracle_envs)
    w, h, k = 512, 1536, 1
    full = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
    acc = np.zeros_like(full)
    for y0, y1 in ((0, 512), (512, 516), (516, 1536)):
        acc |= sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k, rows=(y0, y1))
    assert np.array_equal(acc, full)
    # sharded stripes: two ranks × two stripes
    acc = np.zeros_like(full)
    for r in range(2):
        for y0, y1 in ((0, 768), (768, 1536)):
            acc |= sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k, rows=(y0, y1),
                                     shard_rank=r, shard_count=2, shard_grid=(8, 16), shard_skew=1)
    assert np.array_equal(acc, full)
    # a larger frame through the host-pointer call
    w, h = 2048, 1024
    want = O.render(1, w, h, 0.0, envs=oracle_envs)["fb"]
    got = sr.drawShaderTile(R.FragmentShader.FSDETestShader, None, w, h, 0.0)
    assert np.array_equal(got, want)
    c = sr.counters()
    assert c["last_kernel_ms"] > 0
    with pytest.raises(R.RmdfError):
        sr.drawShaderTile(R.FragmentShader.FSDETestShader, None, 64, 64, 0.0, rows=(2, 64), resolve_log2=1)    # not a multiple of 4
    with pytest.raises(R.RmdfError):
        sr.drawShaderTile(R.FragmentShader.FSDETestShader, 3, 64, 64, 0.0, rows=(0, 32))


def test_fast_mode_within_tolerance(sr, O, R, oracle_envs):
    """Fast mode (FMA contraction, MUFU, HW filtering) vs the oracle at BASELINE's tolerance."""
    _bind(sr, oracle_envs)
    w, h = 512, 512
    ref = O.render(2, w, h, 0.0, envs=oracle_envs)
    got = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, out_format=R.OUT_RGBA32F, mode=R.MODE_FAST)
    agree = np.mean((got[..., 3] > 0.5) == ref["hit"].astype(bool))
    p = psnr(np.clip(np.nan_to_num(got[..., :3]), 0, 1), np.clip(np.nan_to_num(ref["linear"]), 0, 1))
    print("fast mode 512^2 1spp: hit agreement %.5f%%, linear PSNR %.2f dB" % (agree * 100, p))
    assert agree >= 0.999
    assert p >= 30.0        # 1 spp noise floor is ~40 dB; the 50 dB bar is for parity mode


def test_fast_mode_16spp_meets_baseline_tolerance(sr, O, R, oracle_envs):
    """BASELINE's bar at the headline sampling rate (16 spp = 4×4): hit masks >= 99.9 % equal and
    linear-RGB PSNR >= 50 dB on the resolved image, fast build vs the oracle.  A 1024×576 frame
    buffer = 256×144 output pixels keeps the oracle to a second."""
    _bind(sr, oracle_envs)
    w, h, s = 1024, 576, 4
    ref = O.render(2, w, h, 0.0, envs=oracle_envs)
    got = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, out_format=R.OUT_RGBA32F, mode=R.MODE_FAST)
    agree = np.mean((got[..., 3] > 0.5) == ref["hit"].astype(bool))
    box = lambda a: np.clip(np.nan_to_num(a), 0, 1).reshape(h // s, s, w // s, s, 3).mean(axis=(1, 3))
    p = psnr(box(got[..., :3]), box(ref["linear"]))
    print("fast mode 16 spp: hit agreement %.5f%%, linear PSNR %.2f dB" % (agree * 100, p))
    assert agree >= 0.999
    assert p >= 50.0


@pytest.mark.parametrize("scene,flags", [(2, 1), (2, 2), (2, 3), (1, 1), (0, 4), (0, 6), (3, 3)])
def test_shader_options_bit_exact(sr, O, R, oracle_envs, scene, flags):
    """The switches that are off in the reference's build (central-difference normals, CAM_ORTHO,
    RAY_TRACING_TEST): kernel == oracle with the same switch."""
    _bind(sr, oracle_envs)
    w, h = 160, 96
    ref = O.render(scene, w, h, 7.0, envs=oracle_envs, flags=flags)
    got = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, 7.0, flags=flags)
    assert np.array_equal(got, ref["fb"]), int((got != ref["fb"]).sum())
    lin = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, 7.0, flags=flags, out_format=R.OUT_RGBA32F)
    assert np.array_equal(lin[..., 3] > 0.5, ref["hit"].astype(bool))
    assert np.array_equal(lin[..., :3].view(np.uint32), ref["linear"].view(np.uint32))
    fast = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, 7.0, flags=flags, out_format=R.OUT_RGBA32F, mode=R.MODE_FAST)
    assert np.mean((fast[..., 3] > 0.5) == ref["hit"].astype(bool)) > 0.995


@pytest.mark.parametrize("time", [0.0, 3.3, 11.7])
@pytest.mark.parametrize("flags", [0, 1, 2])
def test_cornell_exact_culling(sr, O, R, oracle_envs, time, flags):
    """de_cornell_box skips triangles by three exact tests (supporting plane, bounding sphere, an upper bound
    that starts from the caller's Lipschitz hint and tightens with every evaluated distance): the result must be
    the minimum over all 32 triangles bit for bit — colour, hit mask and work counters equal the oracle's, which
    evaluates every triangle — for several cameras, with central-difference normals (no hint on those calls) and
    with the orthographic camera (parallel rays)."""
    _bind(sr, oracle_envs)
    w, h = 320, 180
    ref = O.render(0, w, h, time, envs=oracle_envs, flags=flags)
    sr.resetCounters()
    got = sr.drawShaderTile(R.FragmentShader.FSDECornellBoxShader, None, w, h, time, flags=flags, out_format=R.OUT_RGBA32F)
    assert np.array_equal(got[..., 3] > 0.5, ref["hit"].astype(bool))
    assert np.array_equal(got[..., :3].view(np.uint32), ref["linear"].view(np.uint32))
    c, rc = sr.counters(), ref["counters"]
    assert (c["samples"], c["in_sphere"], c["hits"], c["march_de"], c["shade_de"]) == \
           (rc["samples"], rc["in_sphere"], rc["hits"], rc["march_de"], rc["shade_de"])
    fast = sr.drawShaderTile(R.FragmentShader.FSDECornellBoxShader, None, w, h, time, flags=flags, out_format=R.OUT_RGBA32F, mode=R.MODE_FAST)
    agree = np.mean((fast[..., 3] > 0.5) == ref["hit"].astype(bool))
    p = psnr(np.clip(np.nan_to_num(fast[..., :3]), 0, 1), np.clip(np.nan_to_num(ref["linear"]), 0, 1))
    print("cornell fast t=%g flags=%d: hit agreement %.5f%%, linear PSNR %.2f dB" % (time, flags, agree * 100, p))
    assert agree >= 0.999 and p >= 35.0


@pytest.mark.parametrize("scene", [0, 1, 2])
def test_soft_shadows_opt_in_not_parity(sr, O, R, oracle_envs, scene):
    """RMDF_FLAG_SOFT_SHADOWS: NOT in the reference (fragment.shd has no shadow code); north_star names
    soft shadows, so they are an opt-in, default off, excluded from parity.  Checked here: without the flag nothing
    changes (the default path equals the oracle), with it the hit mask and every miss pixel stay what they were, hit
    pixels only get darker (some strictly), and the result is deterministic."""
    _bind(sr, oracle_envs)
    w, h = 256, 160
    ref = O.render(scene, w, h, 0.0, envs=oracle_envs)
    base = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, 0.0, out_format=R.OUT_RGBA32F)
    assert np.array_equal(base[..., :3].view(np.uint32), ref["linear"].view(np.uint32))
    for mode in (R.MODE_PARITY, R.MODE_FAST):
        plain = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, 0.0, out_format=R.OUT_RGBA32F, mode=mode)
        sh = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, 0.0, out_format=R.OUT_RGBA32F, mode=mode, flags=R.FLAG_SOFT_SHADOWS)
        again = sr.drawShaderTile(R.FragmentShader(scene), None, w, h, 0.0, out_format=R.OUT_RGBA32F, mode=mode, flags=R.FLAG_SOFT_SHADOWS)
        assert np.array_equal(sh.view(np.uint32), again.view(np.uint32))
        hit = plain[..., 3] > 0.5
        assert np.array_equal(sh[..., 3] > 0.5, hit)
        assert np.array_equal(sh[~hit].view(np.uint32), plain[~hit].view(np.uint32))
        a, b = np.nan_to_num(sh[hit][:, :3]), np.nan_to_num(plain[hit][:, :3])
        assert (a <= b * 1.000001 + 1e-7).all()
        assert (a < b * 0.99).mean() > 0.02, "no pixel is in shadow"


def test_error_behaviour(sr, R):
    with pytest.raises(R.RmdfError):
        sr.drawShaderTile(5, None, 64, 64, 0.0)
    with pytest.raises(R.RmdfError):
        sr.drawShaderTile(R.FragmentShader.FSDETestShader, None, 66, 64, 0.0, resolve_log2=2)
    with pytest.raises(R.RmdfError):
        sr.drawShaderTile(R.FragmentShader.FSDETestShader, None, 64, 64, 0.0, flags=64)
