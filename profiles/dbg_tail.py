"""Tuning variant `tail` (build.py --variant tail --extra=-DRMDF_DEBUG_TAIL=1): per-warp start/end times of one config-4 frame.
   RMDF_LIB=.../librmdf_b200_tail.so python profiles/dbg_tail.py [share_N]"""
import ctypes as C, os, sys, json
sys.path.insert(0, ".")
import numpy as np, torch
import rmdf_b200 as R
from bench import build_env_cubes, FB_W, FB_H
with R.ShaderRenderer(0) as sr:
    build_env_cubes(sr)
    p = sr.makeParams(2, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=R.MODE_FAST)
    out = torch.empty(3840 * 2160, dtype=torch.int32, device="cuda")
    for _ in range(3):
        sr.renderDevice(p, out.data_ptr()); sr.sync()
    ms = sr.counters()["last_kernel_ms"]
    buf = np.zeros(3 * 8192, np.uint64)
    rc = sr._lib.rmdf_debug_tail(buf.ctypes.data_as(C.c_void_p), buf.size)
    t = buf.reshape(8192, 3)
    t = t[t[:, 1] > 0]
    t0, t1 = t[:, 0].min(), t[:, 1].max()
    end = (t[:, 1] - t0).astype(np.float64) * 1e-6
    start = (t[:, 0] - t0).astype(np.float64) * 1e-6
    span = float(t1 - t0) * 1e-6
    idle = span - end
    q = lambda a, f: float(np.quantile(a, f))
    print(json.dumps({"kernel_ms": ms, "warps": int(len(t)), "span_ms": span, "start_ms_max": float(start.max()),
                      "end_ms_quantiles": {k: q(end, f) for k, f in (("p1", .01), ("p10", .1), ("p50", .5), ("p90", .9), ("p99", .99))},
                      "mean_idle_at_end_ms": float(idle.mean()), "items_per_warp_mean": float(t[:, 2].mean()),
                      "items_per_warp_min_max": [int(t[:, 2].min()), int(t[:, 2].max())]}))
