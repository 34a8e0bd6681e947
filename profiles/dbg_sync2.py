"""the test_shared_queue_frame_protocol_on_one_gpu sequence with progress prints (debug)"""
import sys, faulthandler
sys.path.insert(0, "."); sys.path.insert(0, "tests")
faulthandler.enable(); faulthandler.dump_traceback_later(50, exit=True)
import numpy as np, torch
import rmdf_b200 as R
from oracle import oracle as O
def say(*a): print(*a, flush=True)
def view(ptr, n):
    class W: pass
    x = W(); x.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}
    return torch.as_tensor(x, device="cuda:0")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w, h, k = 1024, 512, 1
n_out = (w >> k) * (h >> k)
srs = [R.ShaderRenderer(0) for _ in range(world)]
want = srs[0].drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
frame_bytes = (n_out * 4 + 255) & ~255
ptr, _ = srs[0].sharedAlloc(frame_bytes + 256)
ctrl = ptr + frame_bytes
p = srs[0].makeParams(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
for seq in (1, 2, 3):
    view(ptr, n_out).zero_(); torch.cuda.synchronize(); say("seq", seq, "zeroed")
    for r in srs: r.resetCounters()
    for rank in reversed(range(world)):
        srs[rank].frameBegin(ctrl, rank, world, seq, 5000); say(" begin", rank)
        srs[rank].renderDevice(p, ptr); say(" render", rank)
    srs[0].frameEnd(ctrl, world, seq, 5000); say(" end launched")
    say(" status0", srs[0].frameSyncStatus())
    got = view(ptr, n_out).cpu().numpy().view(np.uint32).reshape(h >> k, w >> k)
    say(" equal", bool(np.array_equal(got, want)), [r.frameSyncStatus() for r in srs[1:]], sum(r.counters()["samples"] for r in srs) == w * h)
jw, jh = 2048, 1024
wantj = srs[0].juliaAnimated(jw, jh, None, False, 7.0, resolve_log2=k); say("julia plain ok")
for rank in reversed(range(world)):
    srs[rank].frameBegin(ctrl, rank, world, 4, 5000)
    srs[rank].juliaDevice(jw, jh, False, 7.0, (0, jh), k, ptr); say(" julia launched", rank)
srs[0].frameEnd(ctrl, world, 4, 5000)
say(" julia status", srs[0].frameSyncStatus())
gotj = view(ptr, (jw >> k) * (jh >> k)).cpu().numpy().view(np.uint32).reshape(jh >> k, jw >> k)
say(" julia equal", bool(np.array_equal(gotj, wantj)))
srs[0].frameBegin(ctrl, 0, 1, 5, 1000)
try:
    srs[0].renderDevice(srs[0].makeParams(R.FragmentShader.FSMBPower8Shader, 3, w, h, 0.0), ptr); say("no error?!")
except R.RmdfError as e:
    say(" expected error:", e)
srs[0].renderDevice(p, ptr); srs[0].frameEnd(ctrl, 1, 5, 1000); say(" status", srs[0].frameSyncStatus())
srs[0].frameEnd(ctrl, 2, 99, 50); say(" timeout status", srs[0].frameSyncStatus())
say("done")
