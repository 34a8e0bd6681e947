"""bench_configs.py — the BASELINE.json configurations that are not the headline (configs[0], [1], [2], [4])
plus the fourth reference scene (general-power Mandelbulb), measured inside bench.py's default N = 1 run and
reported as the `configs` object of the JSON line: kernel time (CUDA events around the launch, on the launching
stream, L2 flushed between launches), the roofline that bounds the kernel, and a bounded CPU baseline.

  roofline.bound  fp32_fma  algorithmic FP32 flop (device counters × SURVEY.md §8d constants) ÷ kernel time,
                            against the in-run FFMA-chain peak
                  xu        algorithmic transcendental calls (sqrt, inversesqrt, log, exp, pow, sin, cos, acos,
                            atan as the shader writes them) ÷ kernel time, against the in-run MUFU peak
                            (16 lanes/clk/SM) — the general-power Mandelbulb and the CSG scene

Only the cpu_baseline legs touch oracle/ (as in bench.py)."""
from __future__ import annotations

import time

import numpy as np

# SURVEY.md §8d constants (GLSL-level counts)
F_ITER8, F_DE_TAIL, F_STEP, F_SAMPLE, F_HIT = 76.0, 9.0, 9.0, 55.0, 130.0
F_TRI, F_CSG_DE = 150.0, 105.0            # per point–triangle distance (×32 per DE); per CSG DE
SFU_CSG_DE = 25.0                         # 10 exp + 5 log + 10 sqrt per CSG DE
SFU_GEN_ITER, SFU_DE_TAIL = 10.0, 1.0     # 2 sqrt + acos + atan + 2 pow + 2 sin + 2 cos per iteration; log per DE
F_GEN_ITER = 30.0                         # flop of the general triplex power + derivative outside the transcendentals
F_JULIA_ITER = 13.0
F_CONV_PAIR = 13.0


def _median_kernel_ms(sr, torch, flush, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_(); fn(); torch.cuda.synchronize()
        ms.append(sr.counters()["last_kernel_ms"])
    return float(np.median(ms))


def _frame_counters(sr, torch, fn):
    sr.resetCounters(); fn(); torch.cuda.synchronize()
    c = sr.counters()
    sr.resetCounters()
    return c


def _march_flops(scene: int, c: dict) -> float:
    de = c["march_de"] + c["shade_de"]
    base = c["march_de"] * F_STEP + c["samples"] * F_SAMPLE + c["hits"] * F_HIT
    if scene == 2:
        return c["inner_iters"] * F_ITER8 + de * F_DE_TAIL + base
    if scene == 3:
        return c["inner_iters"] * F_GEN_ITER + de * F_DE_TAIL + base
    if scene == 0:
        return de * 32.0 * F_TRI + base
    return de * F_CSG_DE + base


def _march_sfu(scene: int, c: dict) -> float:
    de = c["march_de"] + c["shade_de"]
    if scene == 3:
        return c["inner_iters"] * SFU_GEN_ITER + de * SFU_DE_TAIL
    if scene == 1:
        return de * SFU_CSG_DE
    return 0.0


def _psnr_fast_vs_parity(sr, torch, R, scene, fb_w, fb_h, k, time_):
    """level-0 FP32 frames of both builds, s×s box in linear RGB clamped to [0,1] (the resolved image of BASELINE's bar)"""
    s = 1 << k
    imgs = []
    for mode in (R.MODE_PARITY, R.MODE_FAST):
        p = sr.makeParams(R.FragmentShader(scene), None, fb_w, fb_h, time_, out_format=R.OUT_RGBA32F, mode=mode)
        out = torch.empty(fb_w * fb_h * 4, dtype=torch.float32, device="cuda")
        sr.renderDevice(p, out.data_ptr()); sr.sync()
        t = out.reshape(fb_h, fb_w, 4)
        lin = torch.nan_to_num(t[..., :3]).clamp(0, 1).reshape(fb_h // s, s, fb_w // s, s, 3).mean(dim=(1, 3))
        imgs.append((lin.double(), t[..., 3] > 0.5))
        del out, t
    se = float(((imgs[0][0] - imgs[1][0]) ** 2).sum()); n = imgs[0][0].numel()
    agree = float((imgs[0][1] == imgs[1][1]).double().mean())
    return {"hit_agreement": agree, "linear_psnr_db": float(10.0 * np.log10(n / se)) if se > 0 else float("inf"),
            "against": "parity build (bit-identical to the oracle and to the reference's shader compiled by g++)", "spp": s * s,
            "pass": bool(agree >= 0.999 and (se == 0 or 10.0 * np.log10(n / se) >= 50.0))}


def _march_entry(sr, torch, R, flush, scene, fb_w, fb_h, k, time_, peak_tf, mufu_gops, out_dev, name):
    e = {"scene": R.FragmentShader(scene).name, "fb": [fb_w, fb_h], "spp": 1 << (2 * k), "in_time": time_}
    for mode, mname in ((R.MODE_PARITY, "parity"), (R.MODE_FAST, "fast")):
        p = sr.makeParams(R.FragmentShader(scene), None, fb_w, fb_h, time_, resolve_log2=k, mode=mode)
        fn = lambda: sr.renderDevice(p, out_dev.data_ptr())
        c = _frame_counters(sr, torch, fn)
        ms = _median_kernel_ms(sr, torch, flush, fn)
        flops = _march_flops(scene, c)
        sfu = _march_sfu(scene, c)
        m = {"kernel_ms": ms, "samples_per_s": fb_w * fb_h / (ms * 1e-3), "sdf_steps_per_s": (c["march_de"] + c["shade_de"]) / (ms * 1e-3)}
        if scene in (1, 3):
            ach = sfu / (ms * 1e-3) / 1e9
            m["roofline"] = {"bound": "xu", "achieved": ach, "peak": mufu_gops, "unit": "Gop/s", "frac": ach / mufu_gops,
                             "fp32_tflops": flops / (ms * 1e-3) / 1e12, "fp32_frac": flops / (ms * 1e-3) / 1e12 / peak_tf}
        else:
            ach = flops / (ms * 1e-3) / 1e12
            m["roofline"] = {"bound": "fp32_fma", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf}
            if scene == 0:
                m["roofline"]["note"] = ("numerator = the shader's loop over all 32 triangles (150 flop each); the kernel's exact culling "
                                         "evaluates a few of them, so this is a useful-work rate and can exceed the datapath fraction")
        e[mname] = m
    e["tolerance"] = _psnr_fast_vs_parity(sr, torch, R, scene, fb_w, fb_h, k, time_)
    e["conforming_mode"] = "fast" if e["tolerance"]["pass"] else "parity"
    return e


def _cpu_march(scene, fb_w, fb_h, time_, envs, budget_s):
    """bounded CPU baseline for a march config: crops on a 8×6 grid over the frame, all host threads"""
    from bench import cpu_backend, cpu_kind_note
    O, B, kind = cpu_backend()
    cores = O.host_threads()
    pol = O.LOD_ALWAYS_LINEAR if kind == "reference" else O.LOD_QUAD

    def run(cw, ch):
        n, t0 = 0, time.perf_counter()
        for gy in range(6):
            for gx in range(8):
                x0 = ((fb_w // 8) * gx + fb_w // 16 - cw // 2) & ~1
                y0 = ((fb_h // 6) * gy + fb_h // 12 - ch // 2) & ~1
                B.render(scene, fb_w, fb_h, time_, envs=envs, rect=(x0, y0, x0 + cw, y0 + ch), want_linear=False, threads=cores, lod_policy=pol)
                n += cw * ch
        return n, time.perf_counter() - t0
    cw = min(32, fb_w // 8) & ~1
    n, dt = run(cw, cw)
    scale = (budget_s / max(dt, 1e-3)) ** 0.5
    cw2 = int(min(fb_w // 8, max(cw, cw * scale))) & ~1
    ch2 = int(min(fb_h // 6, max(cw, cw * scale))) & ~1
    n, dt = run(cw2, ch2)
    return {"value": n / dt, "unit": "samples/s", "cores": cores, "kind": kind,
            "sample": "48 crops of %dx%d samples on an 8x6 grid over the %dx%d frame buffer (%.1f%% of it); %s" % (cw2, ch2, fb_w, fb_h, 100.0 * n / (fb_w * fb_h), cpu_kind_note(kind))}


def measure(sr, cubes, torch, R, flush, local, stream, with_cpu=True, cpu_budget_s=2.5):
    """Returns the `configs` object.  sr: renderer with the five uffizi cube maps bound."""
    peak_tf = sr.measureFp32Peak()["tflops"]
    mufu = sr.measureMufuPeak()
    out = {"peaks": {"fp32_fma_tflops": peak_tf, "mufu_gops": mufu,
                     "source": "in-run microbenchmarks (csrc/peak.cu): FFMA chains; ex2.approx chains (16 lanes/clk/SM)"}}
    envs = None
    if with_cpu:
        from bench import cpu_backend, oracle_env_cubes
        O, B, kind = cpu_backend()
        envs = oracle_env_cubes(O)
    out_dev = torch.empty(7680 * 4320, dtype=torch.int32, device="cuda")

    # ---- configs[0]: Mandelbulb power 8, 512×512, 1 spp, no env lighting (all five maps unbound = constant 1) ----
    with R.ShaderRenderer(local) as sw:
        sw.setStream(stream.cuda_stream)
        e = _march_entry(sw, torch, R, flush, 2, 512, 512, 0, 0.0, peak_tf, mufu, out_dev, "config1")
        # end to end through the host-pointer call (pageable numpy frame buffer)
        t0 = time.perf_counter()
        for _ in range(20):
            sw.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, 512, 512, 0.0, mode=R.MODE_PARITY)
        e["e2e_ms_host_call_parity"] = (time.perf_counter() - t0) / 20 * 1e3
    e["workload"] = "Mandelbulb power-8 512x512 1 spp, no env lighting"
    if with_cpu:
        e["cpu_baseline"] = _cpu_march(2, 512, 512, 0.0, None, cpu_budget_s)
    out["config1_mandelbulb_512"] = e

    # ---- configs[1]: env-map prefilter of uffizi_512 (256×128, powers 1, 8, 64, 512) ---------------------------
    import ctypes as C
    refl = R.loadHDRImage(__import__("pathlib").Path(__file__).resolve().parent / "data" / "uffizi_512.hdr")
    small = sr.resizeHDRImage(refl, 256)
    h, w, _ = small.shape
    src = torch.from_numpy(small).cuda()
    dst = torch.empty(4 * h * w * 3, dtype=torch.float32, device="cuda")
    pw = (C.c_float * 4)(1.0, 8.0, 64.0, 512.0)
    fn = lambda: R._capi.check(sr._lib.rmdf_envmap_convolve_device(sr.handle, src.data_ptr(), w, h, pw, 4, dst.data_ptr()))
    ms = _median_kernel_ms(sr, torch, flush, fn, reps=7, warm=3)
    pairs = 4.0 * float(w * h) ** 2
    t0 = time.perf_counter()
    for _ in range(5):
        sr.cosineConvolveHDREnvMapSweep(small, sr.powers)
    e2e_ms = (time.perf_counter() - t0) / 5 * 1e3
    ach = pairs * F_CONV_PAIR / (ms * 1e-3) / 1e12
    e = {"workload": "cosine-power convolution of uffizi_512 at 256x128, powers {1, 8, 64, 512} in one sweep", "kernel_ms": ms,
         "pairs_per_s": pairs / (ms * 1e-3), "e2e_ms_host_call": e2e_ms,
         "roofline": {"bound": "fp32_fma", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                      "note": "13 flop per (dst, src) pair and power as the reference writes it; the pow itself (one per pair and power in the reference, nine squarings shared by the four powers here) is not counted"}}
    if with_cpu:
        from oracle import oracle as O2
        cores = O2.host_threads()
        osmall = O2.resize(O2.hdr_decode((__import__("pathlib").Path(__file__).resolve().parent / "data" / "uffizi_512.hdr").read_bytes()), 256)
        rows4 = (62, 64)                                              # 2 of 128 destination rows per power
        t0 = time.perf_counter(); O2.convolve_sweep_rows(osmall, (1.0, 8.0, 64.0, 512.0), rows4); dt4 = time.perf_counter() - t0
        p4 = 4.0 * (rows4[1] - rows4[0]) * w * float(w * h)
        rows_all = (56, 56 + max(cores, 8))
        t0 = time.perf_counter()
        for pwr in (1.0, 8.0, 64.0, 512.0):
            O2.convolve_rows(osmall, pwr, rows_all, threads=cores)
        dta = time.perf_counter() - t0
        pa = 4.0 * (rows_all[1] - rows_all[0]) * w * float(w * h)
        e["cpu_baseline"] = {"value": p4 / dt4, "unit": "pairs/s", "cores": 4, "kind": "port",
                             "sample": "rows %d-%d of the 128 destination rows of each power, one thread per power as the reference runs the sweep (ShaderRendering.hs:142 mapConcurrently; each cosineConvolveHDREnvMap sequential) — C restatement of HDREnvMap.hs:217-254 with libm powf, not GHC" % rows4,
                             "seconds_per_sweep_extrapolated": pairs / (p4 / dt4)}
        e["cpu_baseline_all_cores"] = {"value": pa / dta, "unit": "pairs/s", "cores": cores, "kind": "port",
                                       "sample": "rows %d-%d of each power, OpenMP over destination rows on all host threads" % rows_all,
                                       "seconds_per_sweep_extrapolated": pairs / (pa / dta)}
    out["config2_prefilter"] = e
    del src, dst

    # ---- configs[2]: Cornell box and CSG scene, 1920×1080 × 4 spp (frame buffer 3840×2160), env-lit ------------------
    for scene, key in ((0, "config3_cornell_1080p_4spp"), (1, "config3_csg_1080p_4spp")):
        e = _march_entry(sr, torch, R, flush, scene, 3840, 2160, 1, 0.0, peak_tf, mufu, out_dev, key)
        e["workload"] = "%s 1920x1080 x4spp (FB 3840x2160), uffizi env maps; soft shadows do not exist in the reference (SURVEY D2)" % ("Cornell box" if scene == 0 else "CSG test scene")
        if with_cpu:
            e["cpu_baseline"] = _cpu_march(scene, 3840, 2160, 0.0, envs, cpu_budget_s)
        out[key] = e

    # ---- the fourth reference scene: general-power Mandelbulb (animated power), 3840×2160, 1 spp ------------------------
    e = _march_entry(sr, torch, R, flush, 3, 3840, 2160, 0, 7.0, peak_tf, mufu, out_dev, "general")
    e["workload"] = "general-power Mandelbulb (FSMBGeneralShader, in_time 7: power 5.5) 3840x2160 1 spp, uffizi env maps"
    if with_cpu:
        e["cpu_baseline"] = _cpu_march(3, 3840, 2160, 7.0, envs, cpu_budget_s)
    out["mandelbulb_general_4k"] = e

    # ---- configs[4]: Julia set 7680×4320 × 64 spp (frame buffer 61440×34560) -------------------------------------------
    jw, jh, jk, tick = 61440, 34560, 3, 7.0
    e = {"workload": "Julia set 7680x4320 x64spp (FB 61440x34560), juliaAnimated tick 7", "fb": [jw, jh], "spp": 64}
    for smooth, name in ((False, "plain"), (True, "smooth")):
        fn = lambda: sr.juliaDevice(jw, jh, smooth, tick, (0, jh), jk, out_dev.data_ptr())
        c = _frame_counters(sr, torch, fn)
        ms = _median_kernel_ms(sr, torch, flush, fn, reps=5, warm=2)
        ach = c["inner_iters"] * F_JULIA_ITER / (ms * 1e-3) / 1e12
        e[name] = {"kernel_ms": ms, "samples_per_s": float(jw) * jh / (ms * 1e-3), "iterations": c["inner_iters"],
                   "roofline": {"bound": "fp32_fma", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf}}
    if with_cpu:
        from bench import julia_cpu_baseline
        b = julia_cpu_baseline(budget_s=cpu_budget_s)
        e["cpu_baseline"] = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
    out["config5_julia_8k_64spp"] = e
    del out_dev
    return out
