"""Measures every BASELINE.json config on one GPU (kernel time from the library's CUDA events,
3 warm-ups, best of 5) and writes profiles/<tag>_measurements.json.  Run on the GPU box:

    python profiles/measure_all.py r1
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
import rmdf_b200 as R  # noqa: E402
from bench import build_env_cubes, algorithmic_flops  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
res = {"gpu": torch.cuda.get_device_name(0)}


def kernel_ms(fn, sr, n=5, warm=3):
    for _ in range(warm):
        fn(); sr.sync()
    ms = []
    for _ in range(n):
        fn(); sr.sync()
        ms.append(sr.counters()["last_kernel_ms"])
    return float(min(ms)), float(np.mean(ms))


with R.ShaderRenderer(0) as sr:
    peak = sr.measureFp32Peak()["tflops"]
    res["fp32_peak_tflops"] = peak
    # config 2: prefilter (resize + convolution sweep + cube maps), host pointers in and out
    refl = R.loadHDRImage("data/uffizi_512.hdr")
    t0 = time.perf_counter(); small = sr.resizeHDRImage(refl, 256); t1 = time.perf_counter()
    dsmall = torch.from_numpy(small).cuda(); dout = torch.empty((4,) + small.shape, device="cuda")
    pw = (R._capi.C.c_float * 4)(1, 8, 64, 512)
    conv = lambda: R._capi.check(sr._lib.rmdf_envmap_convolve_device(sr.handle, dsmall.data_ptr(), 256, 128, pw, 4, dout.data_ptr()))
    best, mean = kernel_ms(conv, sr)
    pairs = 4 * (256 * 128) ** 2
    res["config2_envmap_prefilter"] = {"convolve_sweep_ms": best, "pairs": pairs, "gpairs_per_s": pairs / best / 1e6,
                                       "instr_per_pair_model": 28, "resize_e2e_ms": (t1 - t0) * 1e3,
                                       "note": "4 powers {1,8,64,512} of the 256x128 downsample in one pass; time = convolve + finish kernels (CUDA events)"}
    build_env_cubes(sr)

    def frame(scene, w, h, k, mode):
        p = sr.makeParams(scene, None, w, h, 0.0, resolve_log2=k, mode=mode)
        out = torch.empty(sr.outputPixels(p)[0], dtype=torch.int32, device="cuda")
        sr.resetCounters()
        best, mean = kernel_ms(lambda: sr.renderDevice(p, out.data_ptr()), sr)
        c = sr.counters()
        per = {kk: c[kk] / 8 for kk in ("samples", "in_sphere", "hits", "march_de", "shade_de", "inner_iters")}
        d = {"kernel_ms": best, "samples_per_s": w * h / best * 1e3, "sdf_steps_per_s": (per["march_de"] + per["shade_de"]) / best * 1e3,
             "counters_per_frame": {kk: int(v) for kk, v in per.items()}}
        if scene in (2,):
            fl = algorithmic_flops(per)
            d["algorithmic_tflops"] = fl / best / 1e9
            d["frac_fp32_peak"] = d["algorithmic_tflops"] / peak
        return d

    for name, scene, w, h, k in (("config1_mandelbulb_512_1spp", 2, 512, 512, 0),
                                 ("config3_cornell_1080p_4spp", 0, 3840, 2160, 1),
                                 ("config3_csg_1080p_4spp", 1, 3840, 2160, 1),
                                 ("config4_mandelbulb_4k_16spp", 2, 15360, 8640, 2),
                                 ("extra_mandelbulb_general_1080p_4spp", 3, 3840, 2160, 1)):
        if name.startswith("config1"):
            for s in range(5):
                sr.unbindCubeMap(s)
        res[name] = {"fb": [w, h], "resolve_log2": k,
                     "parity": frame(scene, w, h, k, R.MODE_PARITY), "fast": frame(scene, w, h, k, R.MODE_FAST)}
        if name.startswith("config1"):
            build_env_cubes(sr)
    # config 5: Julia 7680x4320 x 64 spp
    w, h, k = 61440, 34560, 3
    out = torch.empty((w >> k) * (h >> k), dtype=torch.int32, device="cuda")
    for smooth in (0, 1):
        sr.resetCounters()
        best, mean = kernel_ms(lambda: sr.juliaDevice(w, h, smooth, 7.0, (0, h), k, out.data_ptr()), sr)
        c = sr.counters()
        res["config5_julia_8k_64spp_smooth%d" % smooth] = {"fb": [w, h], "kernel_ms": best, "samples_per_s": w * h / best * 1e3,
                                                           "iterations_per_frame": c["inner_iters"] // 8,
                                                           "algorithmic_tflops": c["inner_iters"] / 8 * 13 / best / 1e9}
json.dump(res, open("profiles/%s_measurements.json" % tag, "w"), indent=1)
print(json.dumps(res, indent=1))
