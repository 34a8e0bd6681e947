"""Three frames each of config 4 (fast), one rank's eighth of it, and config 1 — for an ncu launch list that shows
order_sort_kernel next to the march kernel:  ncu --metrics gpu__time_duration.sum --clock-control none --csv ..."""
import sys
sys.path.insert(0, ".")
import torch
import rmdf_b200 as R
from bench import build_env_cubes, FB_W, FB_H
sr = R.ShaderRenderer(0)
build_env_cubes(sr)
n_out = 3840 * 2160
frame_bytes = (n_out * 4 + 255) & ~255
ptr, _ = sr.sharedAlloc(frame_bytes + 256)
ctrl = ptr + frame_bytes
p4 = sr.makeParams(2, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=R.MODE_FAST)
p1 = sr.makeParams(2, None, 512, 512, 0.0, mode=R.MODE_FAST)
for i in range(3):
    sr.renderDevice(p4, ptr); sr.sync()
for i in range(3):
    sr.frameBegin(ctrl, 0, 8, i + 1, 50, 0); sr.renderDevice(p4, ptr); sr.sync()
for i in range(3):
    sr.renderDevice(p1, ptr); sr.sync()
