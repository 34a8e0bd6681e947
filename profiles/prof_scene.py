"""One frame per launch of any scene/size, for ncu: python profiles/prof_scene.py <fast|parity> <scene> <w> <h> <k> [n]"""
import sys
sys.path.insert(0, ".")
import torch
import rmdf_b200 as R
from bench import build_env_cubes

mode = R.MODE_FAST if sys.argv[1] == "fast" else R.MODE_PARITY
scene, w, h, k = (int(x) for x in sys.argv[2:6])
n = int(sys.argv[6]) if len(sys.argv) > 6 else 3
with R.ShaderRenderer(0) as sr:
    build_env_cubes(sr)
    p = sr.makeParams(scene, None, w, h, 0.0, resolve_log2=k, mode=mode)
    out = torch.empty(sr.outputPixels(p)[0], dtype=torch.int32, device="cuda")
    for _ in range(n):
        sr.renderDevice(p, out.data_ptr()); sr.sync()
    print("kernel ms", sr.counters()["last_kernel_ms"])
