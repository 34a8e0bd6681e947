"""How even the interleaved multi-GPU deal is: every rank's share of the config-4 frame (and the Julia frame) timed
alone on ONE GPU, with the rank's slot fixed (RMDF_NO_DEAL_ROTATE=1: vertical stripes) and rotated per window.
python profiles/exp_deal.py [WORLD ...]"""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import rmdf_b200 as R
from bench import build_env_cubes, FB_W, FB_H
sr = R.ShaderRenderer(0)
build_env_cubes(sr)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
n_out = 7680 * 4320
frame_bytes = (n_out * 4 + 255) & ~255
ptr, _ = sr.sharedAlloc(frame_bytes + 256)
ctrl = ptr + frame_bytes
p = sr.makeParams(2, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=R.MODE_FAST)
seq = [0]
def share(rank, world, julia):
    seq[0] += 1
    if rank != 0:                                        # publish "go" ourselves: begin as rank 0 on the same context first
        sr.frameBegin(ctrl, 0, world, seq[0], 50, 0); sr.sync()
    sr.frameBegin(ctrl, rank, world, seq[0], 50, 0)
    if julia:
        sr.juliaDevice(61440, 34560, False, 7.0, (0, 34560), 3, ptr)
    else:
        sr.renderDevice(p, ptr)
def t(fn, n=5, warm=3):
    k = []
    for i in range(warm + n):
        flush.zero_(); torch.cuda.synchronize(); fn(); sr.sync()
        if i >= warm:
            k.append(sr.counters()["last_kernel_ms"])
    return float(np.median(k))
for world in [int(a) for a in sys.argv[1:]] or [8]:
    for julia in (False, True):
        for rotate in (False, True):
            if rotate: os.environ.pop("RMDF_NO_DEAL_ROTATE", None)
            else: os.environ["RMDF_NO_DEAL_ROTATE"] = "1"
            ms = [round(t(lambda: share(r, world, julia)), 4) for r in range(world)]
            print(json.dumps({"what": "julia" if julia else "c4_fast", "world": world, "rotate": rotate, "rank_ms": ms,
                              "max": max(ms), "mean": round(float(np.mean(ms)), 4), "max_over_mean": round(max(ms) / float(np.mean(ms)), 4)}), flush=True)
