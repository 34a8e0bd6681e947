"""Times tuning variants of librmdf_b200 (built with `python ray-marching-distance-fields_b200/build.py --variant NAME
--extra "-D..."`) on the GPU box, one subprocess per library (RMDF_LIB selects it):

    python profiles/exp_variants.py [--psnr] [--env K=V ...] NAME [NAME ...]      # NAME = variant name, "default" = the shipped .so

Per variant: kernel ms (library CUDA events, L2 flushed, median of 7) of BASELINE config 4 (fast, parity) and config 1
(fast, parity), optionally the fast-vs-parity PSNR on config 4's centre band.  One JSON line each."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "ray-marching-distance-fields_b200"


def child():
    sys.path.insert(0, str(ROOT))
    import numpy as np
    import torch
    import rmdf_b200 as R
    from bench import build_env_cubes, tolerance_check, FB_W, FB_H
    res = {"variant": os.environ.get("RMDF_VARIANT", "?")}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = torch.empty(3840 * 2160, dtype=torch.int32, device="cuda")

    def ms(sr, p, n=7, warm=2):
        v = []
        for i in range(warm + n):
            flush.zero_(); torch.cuda.synchronize()
            sr.renderDevice(p, out.data_ptr()); sr.sync()
            if i >= warm:
                v.append(sr.counters()["last_kernel_ms"])
        return round(float(np.median(v)), 4), round(float(min(v)), 4)

    which = os.environ.get("RMDF_EXP_WHAT", "c4fast,c4parity,c1")
    with R.ShaderRenderer(0) as sr:
        build_env_cubes(sr)
        if "c4fast" in which:
            res["c4_fast_ms"] = ms(sr, sr.makeParams(2, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=R.MODE_FAST))
        if "c4parity" in which:
            res["c4_parity_ms"] = ms(sr, sr.makeParams(2, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=R.MODE_PARITY), n=3, warm=1)
        if "c3" in which:
            for scene, nm in ((0, "cornell"), (1, "csg")):
                res["c3_%s_fast_ms" % nm] = ms(sr, sr.makeParams(scene, None, 3840, 2160, 0.0, resolve_log2=1, mode=R.MODE_FAST))
                res["c3_%s_parity_ms" % nm] = ms(sr, sr.makeParams(scene, None, 3840, 2160, 0.0, resolve_log2=1, mode=R.MODE_PARITY))
        if "gen" in which:
            res["gen_fast_ms"] = ms(sr, sr.makeParams(3, None, 3840, 2160, 7.0, mode=R.MODE_FAST))
            res["gen_parity_ms"] = ms(sr, sr.makeParams(3, None, 3840, 2160, 7.0, mode=R.MODE_PARITY), n=3, warm=1)
        if "--psnr" in sys.argv:
            t = tolerance_check(sr, torch, R)
            res["c4_psnr_db"], res["c4_hit_agree"] = round(t["linear_psnr_db"], 2), round(t["hit_agreement"], 7)
    if "c1" in which:
        with R.ShaderRenderer(0) as sw:
            res["c1_fast_ms"] = ms(sw, sw.makeParams(2, None, 512, 512, 0.0, mode=R.MODE_FAST), n=15, warm=3)
            res["c1_parity_ms"] = ms(sw, sw.makeParams(2, None, 512, 512, 0.0, mode=R.MODE_PARITY), n=15, warm=3)
    if "julia" in which:
        with R.ShaderRenderer(0) as sj:
            jo = torch.empty(7680 * 4320, dtype=torch.int32, device="cuda")
            for smooth in (0, 1):
                v = []
                for i in range(7):
                    flush.zero_(); torch.cuda.synchronize()
                    sj.juliaDevice(61440, 34560, smooth, 7.0, (0, 34560), 3, jo.data_ptr()); sj.sync()
                    if i >= 2:
                        v.append(sj.counters()["last_kernel_ms"])
                res["julia_smooth%d_ms" % smooth] = (round(float(np.median(v)), 4), round(float(min(v)), 4))
    print(json.dumps(res), flush=True)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    envs = {}
    it = iter(sys.argv[1:])
    names = []
    for a in it:
        if a == "--env":
            k, v = next(it).split("=", 1); envs[k] = v
        elif not a.startswith("--"):
            names.append(a)
    for name in names:
        lib = PKG / ("librmdf_b200.so" if name == "default" else "librmdf_b200_%s.so" % name)
        if not lib.exists():
            print(json.dumps({"variant": name, "error": "not built"})); continue
        env = dict(os.environ, RMDF_LIB=str(lib), RMDF_VARIANT=name, RMDF_EXP_CHILD="1", **envs)
        subprocess.run([sys.executable, __file__] + [a for a in sys.argv[1:] if a.startswith("--psnr")], env=env)


if __name__ == "__main__":
    child() if os.environ.get("RMDF_EXP_CHILD") else main()
