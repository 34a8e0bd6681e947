"""One config-4 frame per launch, for ncu: python profiles/prof_one.py [parity|fast] [scene] [n_launches]"""
import sys
sys.path.insert(0, ".")
import torch
import rmdf_b200 as R
from bench import build_env_cubes, FB_W, FB_H

mode = R.MODE_FAST if (len(sys.argv) < 2 or sys.argv[1] == "fast") else R.MODE_PARITY
scene = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
with R.ShaderRenderer(0) as sr:
    build_env_cubes(sr)
    p = sr.makeParams(scene, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=mode)
    out = torch.empty(sr.outputPixels(p)[0], dtype=torch.int32, device="cuda")
    for _ in range(n):
        sr.renderDevice(p, out.data_ptr()); sr.sync()
    print("kernel ms", sr.counters()["last_kernel_ms"])
