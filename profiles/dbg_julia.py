import sys, faulthandler
sys.path.insert(0, ".")
faulthandler.enable(); faulthandler.dump_traceback_later(60, exit=True)
import numpy as np, torch
import rmdf_b200 as R
def say(*a): print(*a, flush=True)
jw, jh, k = 512, 256, 1
n_out = (jw >> k) * (jh >> k)
sr = R.ShaderRenderer(0)
want = sr.juliaAnimated(jw, jh, None, False, 7.0, resolve_log2=k); say("plain ok")
frame_bytes = (n_out * 4 + 255) & ~255
ptr, _ = sr.sharedAlloc(frame_bytes + 256)
ctrl = ptr + frame_bytes
sr.frameBegin(ctrl, 0, 1, 1, 2000); sr._lib.rmdf_sync(sr.handle); say("begin ok")
sr.juliaDevice(jw, jh, False, 7.0, (0, jh), k, ptr); say("launched")
say("sync rc", sr._lib.rmdf_sync(sr.handle), sr._lib.rmdf_last_error())
sr.frameEnd(ctrl, 1, 1, 2000); say("status", sr.frameSyncStatus())
