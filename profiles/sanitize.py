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
    print("sanitize workload done", sr.counters()["launches"], "launches")
