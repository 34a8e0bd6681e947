"""One rank's share of a config-4 frame dealt to `world` ranks (interleaved), timed alone on one GPU:
python profiles/dbg_share.py WORLD  — kernel ms vs the full-frame kernel / WORLD."""
import sys, os, faulthandler
sys.path.insert(0, ".")
faulthandler.enable(); faulthandler.dump_traceback_later(100, exit=True)
import numpy as np, torch
import rmdf_b200 as R
from bench import build_env_cubes, FB_W, FB_H
world = int(sys.argv[1])
sr = R.ShaderRenderer(0)
build_env_cubes(sr)
n_out = 3840 * 2160
frame_bytes = (n_out * 4 + 255) & ~255
ptr, _ = sr.sharedAlloc(frame_bytes + 256)
ctrl = ptr + frame_bytes
p = sr.makeParams(2, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=R.MODE_FAST)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=7):
    v = []
    for i in range(n):
        flush.zero_(); torch.cuda.synchronize(); fn(); sr.sync(); v.append(sr.counters()["last_kernel_ms"])
    return float(np.median(v[2:]))
full = t(lambda: sr.renderDevice(p, ptr))
seq = [0]
res = []
for rank in (0, world // 2, world - 1):
    def share():
        seq[0] += 1
        sr.frameBegin(ctrl, 0 if rank == 0 else rank, world, seq[0], 50, 0) if rank == 0 else None
        if rank != 0:
            # a non-zero rank would wait for "go": publish it ourselves by beginning as rank 0 on the same ctx first
            sr.frameBegin(ctrl, 0, world, seq[0], 50, 0); sr.sync()
            sr.frameBegin(ctrl, rank, world, seq[0], 50, 0)
        sr.renderDevice(p, ptr)
    res.append(round(t(share), 4))
print("world %d env %s: full %.3f ms, /world %.4f, shares of rank 0/mid/last %s" % (world, {k: v for k, v in os.environ.items() if k.startswith("RMDF_") and k != "RMDF_LIB"}, full, full / world, res), flush=True)
