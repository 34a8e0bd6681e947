"""Latency-bound frames (config 1, CSG): kernel ms against the number of CTAs per SM (RMDF_GRID_CTAS_PER_SM), work
order on (third frame on).  python profiles/exp_grid.py"""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import rmdf_b200 as R
from bench import build_env_cubes
sr = R.ShaderRenderer(0)
build_env_cubes(sr)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = torch.empty(3840 * 2160, dtype=torch.int32, device="cuda")
def t(p, n=11, warm=4):
    k = []
    for i in range(warm + n):
        flush.zero_(); torch.cuda.synchronize()
        sr.renderDevice(p, out.data_ptr()); sr.sync()
        if i >= warm:
            k.append(sr.counters()["last_kernel_ms"])
    return round(float(np.median(k)), 4), round(float(min(k)), 4)
cases = {
    "c1_fast": sr.makeParams(2, None, 512, 512, 0.0, mode=R.MODE_FAST),
    "c1_parity": sr.makeParams(2, None, 512, 512, 0.0, mode=R.MODE_PARITY),
    "c3_csg_fast": sr.makeParams(1, None, 3840, 2160, 0.0, resolve_log2=1, mode=R.MODE_FAST),
    "c3_cornell_fast": sr.makeParams(0, None, 3840, 2160, 0.0, resolve_log2=1, mode=R.MODE_FAST),
    "gen_fast": sr.makeParams(3, None, 3840, 2160, 7.0, mode=R.MODE_FAST),
}
for name, p in cases.items():
    for order in (True, False):
        sr.setWorkOrder(order)
        row = {"what": name, "order": order}
        for cap in ("0", "6", "4", "3", "2"):
            os.environ["RMDF_GRID_CTAS_PER_SM"] = cap
            row["ctas_per_sm_%s" % cap] = t(p)
        print(json.dumps(row), flush=True)
