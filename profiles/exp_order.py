"""Work order (frame_sync.cuh) measured with the permutation made on the HOST (RMDF_ORDER_DEBUG=1: the library records
item costs and leaves the order alone; rmdf_debug_order_get/_set move them): what the ordered queue is worth before
any device-side sort exists.  Kernel ms (library events), L2 flushed, median of 9.
    RMDF_ORDER_DEBUG=1 python profiles/exp_order.py [c4] [share] [c3] [c1]"""
import ctypes as C, json, os, sys, faulthandler
sys.path.insert(0, ".")
faulthandler.enable(); faulthandler.dump_traceback_later(500, exit=True)
import numpy as np, torch
import rmdf_b200 as R
from bench import build_env_cubes, FB_W, FB_H

what = sys.argv[1:] or ["c4", "share", "c3"]
sr = R.ShaderRenderer(0)
L = sr._lib
build_env_cubes(sr)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
n_out = 3840 * 2160
frame_bytes = (n_out * 4 + 255) & ~255
ptr, _ = sr.sharedAlloc(frame_bytes + 256)
ctrl = ptr + frame_bytes
jo = torch.empty(7680 * 4320, dtype=torch.int32, device="cuda")
L.rmdf_debug_order_get.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.rmdf_debug_order_set.argtypes = [C.c_void_p, C.c_void_p, C.c_int]


st = torch.cuda.Stream()
sr.setStream(st.cuda_stream)


def t(fn, n=9, warm=2):
    k, e = [], []
    for i in range(warm + n):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(); e1.record(st); sr.sync()
        if i >= warm:
            k.append(sr.counters()["last_kernel_ms"]); e.append(e0.elapsed_time(e1))
    return {"kernel_ms": round(float(np.median(k)), 4), "kernel_min": round(float(min(k)), 4),
            "stream_ms": round(float(np.median(e)), 4), "stream_min": round(float(min(e)), 4)}


def get_rec():
    buf = np.zeros(1 << 24, np.uint32)
    n = L.rmdf_debug_order_get(sr.handle, buf.ctypes.data, buf.size)
    assert n > 0, n
    return buf[:n].copy()


def set_order(o):
    o = np.ascontiguousarray(o, np.uint32)
    assert np.array_equal(np.sort(o), np.arange(o.size, dtype=np.uint32))
    assert L.rmdf_debug_order_set(sr.handle, o.ctypes.data, o.size) == 0


ORDERS = {
    "identity": lambda c: np.arange(c.size),
    "octave_stable": lambda c: np.argsort(-np.floor(np.log2(np.maximum(c, 1))).astype(np.int64), kind="stable"),
    "half_octave_stable": lambda c: np.argsort(-np.floor(2 * np.log2(np.maximum(c, 1))).astype(np.int64), kind="stable"),
    "exact_desc": lambda c: np.argsort(-c.astype(np.int64), kind="stable"),
    "cheap_quarter_last": lambda c: np.argsort((c < np.quantile(c, 0.25)).astype(np.int64), kind="stable"),
    "octave_shuffled": lambda c: np.lexsort((np.random.default_rng(1).random(c.size), -np.floor(np.log2(np.maximum(c, 1))))),
}


def variants(name, fn, orders=("identity", "octave_stable", "half_octave_stable", "exact_desc", "cheap_quarter_last", "octave_shuffled")):
    sr.setWorkOrder(False)
    print(json.dumps({"what": name, "order": "off", **t(fn)}), flush=True)
    sr.setWorkOrder(True)
    fn(); sr.sync()
    rec = get_rec()
    q = np.quantile(rec, [0, 0.1, 0.5, 0.9, 0.99, 1.0])
    print(json.dumps({"what": name, "items": int(rec.size), "cycles_quantiles_0_10_50_90_99_100": [int(x) for x in q],
                      "mean": float(rec.mean())}), flush=True)
    for o in orders:
        set_order(ORDERS[o](rec))
        print(json.dumps({"what": name, "order": o, **t(fn)}), flush=True)
    if os.environ.pop("RMDF_ORDER_DEBUG", None):
        sr.setWorkOrder(True)                       # the library's own order kernel after every frame
        print(json.dumps({"what": name, "order": "device sort", **t(fn)}), flush=True)
        os.environ["RMDF_ORDER_DEBUG"] = "1"
    sr.setWorkOrder(False)
    print(json.dumps({"what": name, "order": "off again", **t(fn)}), flush=True)


if "c4" in what:
    p = sr.makeParams(2, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=R.MODE_FAST)
    variants("c4_fast", lambda: sr.renderDevice(p, ptr))
if "share" in what:
    p = sr.makeParams(2, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=R.MODE_FAST)
    seq = [0]
    for world in (8,):
        def share():
            seq[0] += 1
            sr.frameBegin(ctrl, 0, world, seq[0], 50, 0)
            sr.renderDevice(p, ptr)
        variants("c4_fast_rank0_of_%d" % world, share)
        os.environ["RMDF_ITEM_H"] = "8"
        variants("c4_fast_rank0_of_%d_item_h8" % world, share, orders=("octave_stable", "exact_desc"))
        os.environ.pop("RMDF_ITEM_H", None)
if "c3" in what:
    for scene, nm in ((0, "cornell"), (1, "csg")):
        p3 = sr.makeParams(scene, None, 3840, 2160, 0.0, resolve_log2=1, mode=R.MODE_FAST)
        variants("c3_%s_fast" % nm, lambda: sr.renderDevice(p3, ptr), orders=("identity", "octave_stable", "exact_desc"))
    pg = sr.makeParams(3, None, 3840, 2160, 7.0, mode=R.MODE_FAST)
    variants("gen_fast", lambda: sr.renderDevice(pg, ptr), orders=("identity", "octave_stable", "exact_desc"))
if "c1" in what:          # needs RMDF_ORDER_MIN_ITEMS=1000 in the environment
    p1 = sr.makeParams(2, None, 512, 512, 0.0, mode=R.MODE_FAST)
    variants("c1_fast", lambda: sr.renderDevice(p1, ptr), orders=("identity", "octave_stable", "exact_desc"))
    p1p = sr.makeParams(2, None, 512, 512, 0.0, mode=R.MODE_PARITY)
    variants("c1_parity", lambda: sr.renderDevice(p1p, ptr), orders=("identity", "octave_stable", "exact_desc"))
