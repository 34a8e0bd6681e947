"""Host-side logic and the C-ABI surface, without a GPU: the shared library loads and exports
every symbol include/rmdf.h declares, argument validation and sizing are pure host code,
compute entry points fail loudly without a device, and the multi-rank gather/assemble path
runs under gloo with world_size 2."""
import ctypes as C
import os
import re
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(R):
    L = R._capi.load(build_if_missing=True)
    hdr = (ROOT / "include" / "rmdf.h").read_text()
    declared = sorted(set(re.findall(r"\b(rmdf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), "librmdf_b200.so does not export %s" % name
    assert sorted(R._capi.EXPORTS) == declared
    assert L.rmdf_abi_version() == 1
    # no torch / python types in the signatures: the header is plain C
    out = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", str(ROOT / "include" / "rmdf.h")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    out = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", str(ROOT / "include" / "rmdf_glsl.h")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure path")
def test_ieee_units_have_no_packed_contraction(R):
    """The bit-exact units are compiled -fmad=false, but ptxas contracts the PACKED forms mul.rn.f32x2 +
    add.rn.f32x2 into FFMA2 regardless (csrc/packed_f32.cuh).  Disassemble the built objects: the parity-march unit
    must hold no packed FMA, the Julia unit only its two intended exact ones, while the fast unit must hold them
    (its iteration is built on FFMA2/FMUL2) — a check that the disassembly itself works."""
    import shutil
    R._capi.load(build_if_missing=True)
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    objdir = ROOT / "ray-marching-distance-fields_b200" / "build"

    def ops(name):
        sass = subprocess.run([cuobjdump, "-sass", str(objdir / name)], capture_output=True, text=True, check=True).stdout
        return re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", sass, flags=re.M)

    o = ops("march_parity.cu.o")
    assert len(o) > 1000 and "FFMA2" not in o, "march_parity.cu.o: ptxas contracted a packed multiply-add"
    # julia.cu writes two EXACT packed FMAs on purpose — 2·p + ci and zi²·(−1) + zr² (csrc/julia.cu, RMDF_JULIA_PAIR) — whose
    # multiplier is the literal 2 or −1; a contraction of a product into a sum would have two register multiplicands
    sass = subprocess.run([cuobjdump, "-sass", str(objdir / "julia.cu.o")], capture_output=True, text=True, check=True).stdout
    fma2 = [l for l in sass.splitlines() if re.search(r"\bFFMA2\b", l)]
    assert len(fma2) >= 48
    for l in fma2:
        assert re.search(r"FFMA2\s+R\d+, R\d+(\.\w+)*, (2|-1), ", l), "julia.cu.o: ptxas contracted a packed multiply-add: " + l.strip()
    assert "FMUL2" in ops("julia.cu.o") and "FADD2" in ops("julia.cu.o")
    fast = ops("march_fast.cu.o")
    assert "FFMA2" in fast and "FMUL2" in fast


def test_no_cpu_fallback(R):
    with pytest.raises(R.RmdfError) as e:
        R.ShaderRenderer(0)
    assert e.value.code == R._capi.ERR_CUDA and "no CPU fallback" in str(e.value)


def _params(R, **kw):
    p = R.Params()
    p.scene, p.fb_w, p.fb_h, p.tile_idx, p.shard_count = 2, 256, 128, -1, 1
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def test_output_sizing_and_validation(R):
    L = R._capi.load()
    nt = C.c_int32()
    n = lambda **kw: L.rmdf_render_output_pixels(C.byref(_params(R, **kw)), C.byref(nt))
    assert n() == 256 * 128 and nt.value == 1
    assert n(resolve_log2=2) == 64 * 32
    assert n(tile_idx=5) == 256 * 128 and nt.value == 1                     # frame layout: untouched pixels stay
    assert n(tile_idx=5, layout=R.LAYOUT_PACKED_TILES) == 32 * 16
    assert n(tile_idx=69, layout=R.LAYOUT_PACKED_TILES) == 32 * 16          # idx mod 64
    # any frame size (ShaderRendering.hs:183-195): a pixel belongs to the tile that contains its centre
    edge = lambda t, dim: (2 * t * dim + 7) >> 4
    tot = 0
    for idx in range(64):
        v = n(fb_w=100, fb_h=45, tile_idx=idx, layout=R.LAYOUT_PACKED_TILES)
        assert v == (edge(idx % 8 + 1, 100) - edge(idx % 8, 100)) * (edge(idx // 8 + 1, 45) - edge(idx // 8, 45))
        tot += v
    assert tot == 100 * 45 and n(fb_w=100, fb_h=45, tile_idx=3) == 100 * 45
    for g in (2, 4, 8):
        assert n(shard_count=g, shard_rank=g - 1, layout=R.LAYOUT_PACKED_TILES, resolve_log2=1) == (64 // g) * 16 * 8
        assert nt.value == 64 // g
    # errors: every one sets rmdf_last_error
    for bad in (dict(scene=4), dict(fb_w=0), dict(resolve_log2=4), dict(out_format=2), dict(mode=7), dict(lod_policy=3),
                dict(out_format=R.OUT_RGBA32F, resolve_log2=1), dict(fb_w=250, resolve_log2=2), dict(fb_w=100, fb_h=100, tile_idx=1, resolve_log2=1),
                dict(shard_count=2, shard_rank=2), dict(layout=R.LAYOUT_PACKED_TILES), dict(shard_count=3, shard_rank=0),
                dict(shard_count=2, tile_idx=1), dict(shard_count=2, shard_grid_x=7, shard_grid_y=8), dict(shard_count=2, shard_grid_x=64, shard_grid_y=8, fb_w=320)):
        assert n(**bad) == -1, bad
        assert len(L.rmdf_last_error()) > 0
    assert n(shard_count=4, shard_rank=1, shard_grid_x=32, shard_grid_y=16, shard_skew=2, layout=R.LAYOUT_PACKED_TILES) == 256 * 128 // 4
    assert nt.value == 32 * 16 // 4
    p = _params(R); p.flags = 16
    assert L.rmdf_render_output_pixels(C.byref(p), None) == -1
    assert L.rmdf_render_output_pixels(None, None) == -1


def test_host_only_entry_points_match_the_oracle(R, O, uffizi):
    L = R._capi.load()
    xyz = np.empty((96, 3), np.float32)
    assert L.rmdf_cornell_box_vertices(xyz.ctypes.data) == 0
    assert np.array_equal(xyz, O.cornell_vertices())
    assert abs(np.abs(xyz).max() - 0.99 / np.sqrt(3)) < 1e-6                # fits the R = 1 bounding sphere
    assert L.rmdf_cornell_box_vertices(None) != 0
    img = (uffizi * np.float32(0.731)).copy()
    mine = R.ShaderRenderer.rgbeRoundTrip(img)
    assert np.array_equal(mine, O.rgbe_roundtrip(img))
    assert L.rmdf_envmap_resize_height(512, 256, 256) == 128 == O.lib().oracle_resize_height(512, 256, 256)
    assert L.rmdf_envmap_resize_height(500, 300, 7) == O.lib().oracle_resize_height(500, 300, 7)
    assert L.rmdf_envmap_resize_height(0, 0, 0) == 0


def test_hdr_file_io(R, O, uffizi, tmp_path):
    raw = (ROOT / "data" / "uffizi_512.hdr").read_bytes()
    assert np.array_equal(R.decodeHDR(raw), uffizi)
    assert R.encodeHDR(uffizi) == O.hdr_encode(uffizi)
    # new-style RLE scanlines decode to the same pixels (hand-encode two rows)
    rgbe = np.frombuffer(R.encodeHDR(uffizi[:2, :16]), np.uint8)[-2 * 16 * 4:].reshape(2, 16, 4)
    rle = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 16\n"
    for y in range(2):
        rle += bytes([2, 2, 0, 16])
        for c in range(4):
            rle += bytes([16]) + rgbe[y, :, c].tobytes()
    assert np.array_equal(R.decodeHDR(rle), O.hdr_decode(rle))
    assert np.array_equal(R.decodeHDR(rle), R.decodeHDR(R.encodeHDR(uffizi[:2, :16])))
    with pytest.raises(ValueError):
        R.loadHDRImage(tmp_path / "missing.hdr")
    (tmp_path / "bad.hdr").write_bytes(b"P6\n1 1\n255\n\0\0\0")
    with pytest.raises(ValueError):
        R.loadHDRImage(tmp_path / "bad.hdr")
    # cache-file naming of ShaderRendering.hs:72-75 (`show (8 :: Float)` = "8.0")
    assert R.ShaderRenderer.mkCacheFn("./latlong_envmaps/uffizi_512.hdr", 8.0) == "./latlong_envmaps/uffizi_512_cache_pow_8.0.hdr"
    assert R.ShaderRenderer.mkCacheFn("a/b.hdr", 512).endswith("b_cache_pow_512.0.hdr")


def test_tile_enum_and_framebuffer(R, O, tmp_path):
    assert [int(s) for s in R.FragmentShader] == [0, 1, 2, 3]
    assert R.isTileIdxFirstTile(0) and R.isTileIdxFirstTile(128) and not R.isTileIdxFirstTile(1)
    assert R.isTileIdxLastTile(63) and R.isTileIdxLastTile(127) and not R.isTileIdxLastTile(64)
    fb = R.FrameBuffer(4, 2)
    assert fb.getFrameBufferDim() == (4, 2) and (fb.data == 0xFF000000).all()
    fb.fillFrameBuffer(lambda w, h, v: v.__setitem__(slice(None), np.arange(w * h, dtype=np.uint32) << 8))
    fb.saveFrameBufferToPNG(tmp_path / "s.png")
    from PIL import Image
    img = np.array(Image.open(tmp_path / "s.png"))
    assert np.array_equal(img, O.screenshot_rows(fb.data))                  # flipped rows, alpha 0xFF


def test_tile_dealing_and_assembly(R):
    from importlib import import_module
    par = import_module("ray-marching-distance-fields_b200.parallel")
    for g in (1, 2, 4, 8):
        for grid, skew in (((8, 8), 1), ((32, 16), 2), ((16, 8), 3)):
            gx, gy = grid
            owned = [par.tiles_of_rank(r, g, grid, skew) for r in range(g)]
            assert sorted(sum(owned, [])) == list(range(gx * gy)) and all(len(o) == gx * gy // g for o in owned)
            # every rank holds tiles from every tile row: balanced load
            assert all(len({t // gx for t in o}) == gy for o in owned)
            w, h = 128, 64
            frame = np.arange(w * h, dtype=np.uint32).reshape(h, w)
            tiles = frame.reshape(gy, h // gy, gx, w // gx).transpose(0, 2, 1, 3).reshape(gx * gy, h // gy, w // gx)
            packed = [tiles[o].reshape(-1) for o in owned]
            assert np.array_equal(par.assemble_tiles_numpy(packed, w, h, g, grid, skew), frame)
    assert par.julia_band_height(34560, 3, 8) == 96 and par.julia_band_height(34560, 3, 1) % 32 == 0      # 45 bands per rank
    bands = [np.full((2, 4, 6), r) for r in range(3)]          # 3 ranks × 2 bands × 4 rows × 6 px
    fr = par.assemble_bands_numpy([b.reshape(-1) for b in bands], 6, 24, 3, 4)
    assert fr[:, 0].tolist() == [0] * 4 + [1] * 4 + [2] * 4 + [0] * 4 + [1] * 4 + [2] * 4
    # finest grid of >= 64-sample tiles on even pixel boundaries, gx a multiple of the world size, skew G/2
    assert par.balanced_grid(15360, 8640, 2, 8) == ((240, 135), 4)
    assert par.balanced_grid(15360, 8640, 2, 2) == ((240, 135), 1)
    (gx, gy), skew = par.balanced_grid(1024, 512, 2, 4)
    assert gx % 4 == 0 and 1024 % (2 * gx) == 0 and 512 % (2 * gy) == 0 and (1024 // gx) % 4 == 0 and (512 // gy) % 4 == 0


_GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from importlib import import_module
par = import_module("ray-marching-distance-fields_b200.parallel")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=world)
w, h = 128, 64
frame = (np.arange(w * h, dtype=np.int64) * 2654435761 % (1 << 31)).astype(np.int32).reshape(h, w)
grid, skew = (16, 8), 2
tiles = frame.reshape(8, h // 8, 16, w // 16).transpose(0, 2, 1, 3).reshape(128, -1)
mine = torch.from_numpy(np.ascontiguousarray(tiles[par.tiles_of_rank(rank, world, grid, skew)]).reshape(-1))
out = par.gather_frame(mine, w, h, world, rank, grid=grid, skew=skew)
if rank == 0:
    assert out is not None and np.array_equal(out.numpy(), frame), "assembled frame differs"
else:
    assert out is None
# the shared HOST frame of the end-to-end path: one POSIX shm segment, every rank writes its tiles in place
class FakeRenderer:                       # stands in for the CUDA context: "registering" returns the host address
    def hostRegister(self, ptr, nbytes): return ptr
    def hostUnregister(self, ptr): pass
sh = par.SharedHostFrame(FakeRenderer(), w * h, world, rank)
view = np.frombuffer(sh.shm.buf, dtype=np.int32, count=w * h).reshape(h, w)
tw, th = w // grid[0], h // grid[1]
for t in par.tiles_of_rank(rank, world, grid, skew):
    ty, tx = divmod(t, grid[0])
    view[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw] = frame[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
sh.complete()
if rank == 0:
    assert np.array_equal(sh.as_numpy().view(np.int32).reshape(h, w), frame), "shared host frame differs"
    print("GLOO_OK")
del view
sh.close()
dist.barrier(); dist.destroy_process_group()
'''


def test_gather_frame_gloo_world_size_2(tmp_path):
    """The N > 1 path of bench.py (tile sharding + one gather to rank 0) on CPU."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(ROOT), str(port)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GLOO_OK" in outs[0]


def test_bench_reference_arm_contract():
    """--impl reference prints one JSON line with the contract's keys (tiny budget)."""
    import json
    # OMP_NUM_THREADS=1 is what torch.distributed.run exports when nproc > 1: the CPU legs pass their thread
    # count explicitly, so it must not apply (round 1's reference arm ran single-threaded under torchrun)
    env = dict(os.environ, RMDF_BENCH_CPU_BUDGET_S="0.5", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["value"] > 0
    from oracle import oracle_ref
    assert line["cpu_baseline"]["kind"] == ("reference" if oracle_ref.available() else "port")
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # same workload keys as the GPU arm's config
    for k, v in (("scene", "FSMBPower8Shader"), ("in_time", 0.0), ("fb", [15360, 8640]), ("out", [3840, 2160]), ("spp", 16)):
        assert line["config"][k] == v
    # rank != 0 of a torchrun launch exits 0 without printing
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=600)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_threads_are_real(O):
    """The thread count the CPU legs report is the one OpenMP uses, whatever OMP_NUM_THREADS says: the same crop
    rendered with threads=1 and threads=all must show a speed-up when more than one core is available."""
    import time
    cores = O.host_threads()
    if cores < 4:
        pytest.skip("needs >= 4 host threads")
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        t0 = time.perf_counter(); O.render(2, 512, 512, 0.0, threads=1, want_linear=False); t1 = time.perf_counter() - t0
        t0 = time.perf_counter(); O.render(2, 512, 512, 0.0, threads=cores, want_linear=False); tn = time.perf_counter() - t0
    finally:
        os.environ.pop("OMP_NUM_THREADS", None)
    assert tn < t1 / 1.8, (t1, tn, cores)


def test_committed_bench_lines_follow_the_contract():
    """profiles/r2_bench_n{1,2,8}.json are the raw lines of `bench.py` runs on B200s (kept as evidence): each must be
    the ONE JSON line of the driver's contract — metric, whole-job value, device-timed ms, roofline with its fraction,
    the end-to-end leg with its copy sizes, kernel launches, clocks without thermal throttling — and consistent in itself."""
    import json
    for n in (1, 2, 8):
        p = ROOT / "profiles" / ("r2_bench_n%d.json" % n)
        line = json.loads(p.read_text().strip().splitlines()[-1])
        assert line["metric"] == "mandelbulb_4k_samples_per_s" and line["unit"] == "samples/s" and line["n_gpus"] == n
        assert line["higher_is_better"] is True and line["scaling"] == "strong" and line["dtype"] == "f32" and line["data"] == "synthetic"
        assert line["steps"] >= 10 and line["warmup"] >= 3
        samples = 15360 * 8640
        assert abs(line["value"] - samples / (line["ms_per_step"] * 1e-3)) <= 1e-6 * line["value"]
        assert line["config"]["fb"] == [15360, 8640] and line["config"]["spp"] == 16 and "workload" in line["config"]
        r = line["roofline"]
        assert r["bound"] == "fp32_fma" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert 0.3 < r["frac"] < 1.0 and r["kernel_ms"] <= line["ms_per_step"] * 1.01
        e = line["e2e"]
        assert e["h2d_bytes_per_step"] > 4_000_000 and e["d2h_bytes_per_step"] == 3840 * 2160 * 4
        assert e["value"] < line["value"]                          # copies inside the timed region cost something
        assert line["gpu_launches"] >= line["steps"]
        c = line["clocks"]
        assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert line["tolerance_check"]["pass"] is True and line["tolerance_check"]["linear_psnr_db"] >= 50.0
        w = line["work_order"]
        assert w["enabled"] is True and w["kernel_ms"] > 0 and w["scan_order_kernel_ms"] > 0
        if n == 1:
            assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
            assert line["parity_check"]["bit_identical"] is True and line["parity_check"]["samples_differing"] == 0
            assert set(line["configs"]) >= {"config1_mandelbulb_512", "config2_prefilter", "config3_cornell_1080p_4spp",
                                            "config3_csg_1080p_4spp", "config5_julia_8k_64spp", "mandelbulb_general_4k"}
        else:
            assert line["frame_equals_single_gpu"] is True and len(line["rank_kernel_ms"]) == n
            assert line["julia_8k_64spp"]["frame_equals_single_gpu"] is True
