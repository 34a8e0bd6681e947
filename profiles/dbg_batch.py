"""config 4 fast on ONE GPU: plain local queue vs the shared-queue protocol with world = 1 (system-scope claims, batches)"""
import sys, os, faulthandler
sys.path.insert(0, ".")
faulthandler.enable(); faulthandler.dump_traceback_later(100, exit=True)
import numpy as np, torch
import rmdf_b200 as R
from bench import build_env_cubes, FB_W, FB_H
sr = R.ShaderRenderer(0)
build_env_cubes(sr)
n_out = 3840 * 2160
frame_bytes = (n_out * 4 + 255) & ~255
ptr, _ = sr.sharedAlloc(frame_bytes + 256)
ctrl = ptr + frame_bytes
p = sr.makeParams(2, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=R.MODE_FAST)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=6):
    v = []
    for i in range(n):
        flush.zero_(); torch.cuda.synchronize(); fn(); sr.sync(); v.append(sr.counters()["last_kernel_ms"])
    return round(float(np.median(v[1:])), 3)
print("plain", t(lambda: sr.renderDevice(p, ptr)), flush=True)
seq = [0]
def shared():
    seq[0] += 1
    sr.frameBegin(ctrl, 0, 1, seq[0], 3000); sr.renderDevice(p, ptr); sr.frameEnd(ctrl, 1, seq[0], 3000)
print("shared world=1 batch_max=%s" % os.environ.get("RMDF_BATCH_MAX", "default"), t(shared), flush=True)
