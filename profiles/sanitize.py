"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel once."""
import sys
sys.path.insert(0, ".")
import numpy as np
import rmdf_b200 as R

with R.ShaderRenderer(0) as sr:
    refl = R.loadHDRImage("data/uffizi_512.hdr")
    small = sr.resizeHDRImage(refl, 64)
    conv = sr.cosineConvolveHDREnvMapSweep(small, [1.0, 8.0, 64.0, 512.0])
    sr.cosineConvolveHDREnvMapSweep(small[:8, :16].copy(), [2.5])
    sr.latLongHDREnvMapToCubeMap(0, small)
    for i in range(4):
        sr.latLongHDREnvMapToCubeMap(1 + i, conv[i])
    for scene in range(4):
        for mode in (0, 1):
            sr.drawShaderTile(scene, None, 96, 64, 1.0, mode=mode, resolve_log2=1)
            sr.drawShaderTile(scene, 9, 96, 64, 1.0, mode=mode)
    sr.drawShaderTile(2, None, 67, 45, 1.0)
    sr.drawShaderTile(0, None, 64, 64, 1.0, flags=7 & ~2)
    sr.drawShaderTile(2, None, 128, 64, 1.0, shard_rank=1, shard_count=2, layout=R.LAYOUT_PACKED_TILES, resolve_log2=2)
    sr.juliaAnimated(97, 33, None, True, 3.0)
    sr.juliaAnimated(128, 64, None, False, 3.0, resolve_log2=3)
    # round 2: tiles of a frame whose size is not a multiple of 16 (helper quads, clipped stores), soft shadows,
    # a ragged convolution, and the multi-GPU frame protocol with two contexts on this GPU (both deals, march + Julia)
    for idx in (0, 9, 27, 63):
        sr.drawShaderTile(1, idx, 100, 45, 1.0)
        sr.drawShaderTile(1, idx, 100, 45, 1.0, layout=R.LAYOUT_PACKED_TILES, out_format=R.OUT_RGBA32F)
    sr.drawShaderTile(2, None, 64, 48, 1.0, flags=R.FLAG_SOFT_SHADOWS, mode=1)
    sr.cosineConvolveHDREnvMapSweep(sr.resizeHDRImage(refl, 48), [1.0, 8.0, 64.0, 512.0, 3.0])
    with R.ShaderRenderer(0) as b:
        w, h, k = 128, 64, 1
        n_out = (w >> k) * (h >> k)
        fb = (n_out * 4 + 255) & ~255
        ptr, _ = sr.sharedAlloc(fb + 256)
        p = sr.makeParams(2, None, w, h, 1.0, resolve_log2=k, mode=1)
        seq = 0
        for deal in (0, 1):
            seq += 1
            b.frameBegin(ptr + fb, 1, 2, seq, 3000, deal); b.renderDevice(p, ptr)
            sr.frameBegin(ptr + fb, 0, 2, seq, 3000, deal); sr.renderDevice(p, ptr)
            sr.frameEnd(ptr + fb, 2, seq, 3000)
            assert sr.frameSyncStatus() == 0 and b.frameSyncStatus() == 0
            seq += 1
            b.frameBegin(ptr + fb, 1, 2, seq, 3000, deal); b.juliaDevice(w, h, False, 3.0, (0, h), k, ptr)
            sr.frameBegin(ptr + fb, 0, 2, seq, 3000, deal); sr.juliaDevice(w, h, False, 3.0, (0, h), k, ptr)
            sr.frameEnd(ptr + fb, 2, seq, 3000)
            assert sr.frameSyncStatus() == 0 and b.frameSyncStatus() == 0
        b.sync(); sr.sharedFree(ptr)
    # round 2, third session: the cost-ordered queue (run with RMDF_ORDER_MIN_ITEMS=64 so that these small frames qualify):
    # record + order_place_kernel + ordered claims, whole frames and one rank's share of a rotated two-rank deal
    for mode in (0, 1):
        pf = sr.makeParams(2, None, 256, 128, 1.0, mode=mode)
        ptr, _ = sr.sharedAlloc(256 * 128 * 4 + 512)
        for _ in range(3):
            sr.renderDevice(pf, ptr)
        for seq in (1, 2, 3):
            sr.frameBegin(ptr + 256 * 128 * 4, 0, 2, seq, 3000, 0); sr.renderDevice(pf, ptr)
        sr.sync(); sr.sharedFree(ptr)
    print("sanitize workload done", sr.counters()["launches"], "launches")
