#!/usr/bin/env python
"""bench.py — headline benchmark: BASELINE.json's "Mandelbulb 4K samples/s".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one frame of BASELINE config 4: Mandelbulb power 8 (fragment.shd), frame buffer
15360×8640 = 3840×2160 × 16 spp, lit by the prefiltered uffizi_512 environment maps,
resolved in-kernel to 3840×2160 RGBA8.  Strong scaling: the frame's 8×8 tiles are dealt to
the N ranks and gathered to rank 0 over NCCL.

  value            samples/s with everything resident in HBM (CUDA events, max over ranks)
  e2e              the same frame through the host-pointer C-ABI call: per step H2D of the five
                   cube maps from pinned memory, kernel, (NCCL gather,) D2H of the resolved frame
  roofline         FP32-FMA: algorithmic flop (device counters × SURVEY §8d constants) ÷ kernel
                   time, against an in-run FFMA microbenchmark (MEASURED_PEAKS.json has no FP32
                   number) — plus the HBM side to show the kernel is not bandwidth-bound
  cpu_baseline     the CPU oracle (C restatement of the reference; llvmpipe/GHC are not in the
                   image) on all host cores over a bounded sample of the same frame

--impl reference times that CPU restatement alone (the reference itself — Haskell + GLSL —
cannot be built here: no GHC, no OpenGL).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner at the first
# collective), so file descriptor 1 is pointed at stderr for everything else and the result line goes to the
# original stdout.
# original stdout (done in main(), not on import: profiles/*.py import helpers from this file).
_RESULT_OUT = None


def _claim_stdout() -> None:
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# ---- workload: BASELINE.json configs[3] ---------------------------------------------------
OUT_W, OUT_H, SPP_LOG2 = 3840, 2160, 2                   # 16 spp = 4×4 ordered grid
FB_W, FB_H = OUT_W << SPP_LOG2, OUT_H << SPP_LOG2        # 15360 × 8640 samples
SCENE_MB_POWER8 = 2
TIME = 0.0
WORKLOAD = "Mandelbulb power-8 3840x2160 x16spp (FB 15360x8640), prefiltered uffizi_512 env maps"

# algorithmic FP32 flop constants, SURVEY.md §8d (GLSL-level count; transcendentals not counted)
F_ITER, F_DE_TAIL, F_STEP, F_SAMPLE, F_HIT = 76.0, 9.0, 9.0, 55.0, 130.0


def algorithmic_flops(c: dict) -> float:
    de = c["march_de"] + c["shade_de"]
    return (c["inner_iters"] * F_ITER + de * F_DE_TAIL + c["march_de"] * F_STEP +
            c["samples"] * F_SAMPLE + c["hits"] * F_HIT)


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ---- environment maps (host side, product path: librmdf_b200 only) ------------------------
def build_env_cubes(sr):
    """uffizi_512.hdr → resize 256 → cos^{1,8,64,512} → RGBE → cube maps, all on the GPU through
    the C ABI; returns the five face arrays (host, float32) for per-step uploads."""
    import rmdf_b200 as R
    refl = R.loadHDRImage(ROOT / "data" / "uffizi_512.hdr")
    small = sr.resizeHDRImage(refl, 256)
    conv = sr.cosineConvolveHDREnvMapSweep(small, sr.powers)
    sr.latLongHDREnvMapToCubeMap(R.ENV_REFLECTION, refl)
    for i in range(4):
        sr.latLongHDREnvMapToCubeMap(R.ENV_COS_1 + i, sr.rgbeRoundTrip(conv[i]))
    return [sr.getCubeMap(s) for s in range(5)]


def tolerance_check(sr, torch, R):
    """BASELINE.json's acceptance bar for the fast build, measured in this run against the parity
    build (which tests/test_march_parity.py pins bit-for-bit to the CPU oracle): hit/miss masks agree
    on >= 99.9 % of samples and linear-RGB PSNR >= 50 dB on the 16-spp resolved image.  Region: the
    centre band of the frame (tiles 24-39 of the 8x8 grid, where all the geometry is), level-0 FP32
    output, 4x4 box-averaged in linear space, colours clamped to [0,1] like the frame buffer."""
    s = 1 << SPP_LOG2
    tw, th = FB_W // 8, FB_H // 8
    se, n, agree = 0.0, 0, 0.0
    for idx in range(24, 40):
        imgs = []
        for mode in (R.MODE_PARITY, R.MODE_FAST):
            p = sr.makeParams(R.FragmentShader.FSMBPower8Shader, idx, FB_W, FB_H, TIME, out_format=R.OUT_RGBA32F,
                              mode=mode, layout=R.LAYOUT_PACKED_TILES)
            out = torch.empty(tw * th * 4, dtype=torch.float32, device="cuda")
            sr.renderDevice(p, out.data_ptr()); sr.sync()
            t = out.reshape(th, tw, 4)
            lin = torch.nan_to_num(t[..., :3]).clamp(0, 1).reshape(th // s, s, tw // s, s, 3).mean(dim=(1, 3))
            imgs.append((lin.double(), t[..., 3] > 0.5))
        se += float(((imgs[0][0] - imgs[1][0]) ** 2).sum()); n += imgs[0][0].numel()
        agree += float((imgs[0][1] == imgs[1][1]).double().mean()) / 16.0
    psnr = float(10.0 * np.log10(n / se)) if se > 0 else float("inf")
    return {"against": "parity build (bit-identical to the CPU oracle)", "region": "tiles 24-39 (centre band), level-0 FP32, 4x4 box in linear RGB",
            "hit_agreement": agree, "linear_psnr_db": psnr, "required": {"hit_agreement": 0.999, "linear_psnr_db": 50.0},
            "pass": bool(agree >= 0.999 and psnr >= 50.0)}


class ClockSampler:
    """nvidia-smi sampled every 50 ms while the timed region runs (the profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            return
        t0 = time.perf_counter()                      # nvidia-smi needs a few hundred ms to print its first row
        while time.perf_counter() - t0 < 3.0 and os.path.getsize(self.f.name) == 0 and self.p.poll() is None:
            time.sleep(0.02)

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.06)                              # one more sampling period
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons, power, busy = [], [], set(), [], []
        for line in self.f.read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1])); mx.append(float(t[2])); power.append(float(t[3]))
            except ValueError:
                continue
            try:
                busy.append(float(t[4]) > 0.0)
            except ValueError:
                busy.append(False)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            load = [c for c, b in zip(sm, busy) if b] or sm
            out = {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "power_w_max": float(max(power)), "samples": len(sm), "samples_under_load": int(sum(busy))}
        return out


# ---- CPU baseline (oracle; the only place bench.py touches oracle/) -----------------------
def cpu_backend():
    """The CPU implementation that is timed: oracle/_ref — the reference's OWN fragment.shd compiled by g++
    (kind "reference") — when the prebuilt library is here, else the C restatement (kind "port").
    Both take an explicit thread count (OpenMP num_threads clause), so torchrun's OMP_NUM_THREADS=1
    does not apply."""
    from oracle import oracle as O
    try:
        from oracle import oracle_ref as REF
        if REF.available():
            return O, REF, "reference"
    except Exception:
        pass
    return O, O, "port"


def cpu_kind_note(kind: str) -> str:
    return ("the reference's own fragment.shd compiled by g++ -O2 (oracle/_ref; GLSL built-ins = include/rmdf_glsl.h, LINEAR filtering — "
            "the quad LOD rule would re-run NEAREST quads), OpenMP over rows of 2x2 quads; not llvmpipe (no GL in the image)" if kind == "reference"
            else "C restatement of fragment.shd with OpenMP — not llvmpipe/GHC (absent from the image)")


def oracle_env_cubes(O):
    img = O.hdr_decode((ROOT / "data" / "uffizi_512.hdr").read_bytes())
    small = O.resize(img, 256)
    return [O.latlong_to_cube(img)] + [O.latlong_to_cube(O.rgbe_roundtrip(O.convolve(small, p))) for p in (1.0, 8.0, 64.0, 512.0)]


def cpu_baseline_sample(steps: int = 1, warmup: int = 0, budget_s: float = float(os.environ.get("RMDF_BENCH_CPU_BUDGET_S", "12")), keep=None):
    """Times the CPU implementation on a bounded, frame-covering sample of config 4: a 16×9 grid of
    square crops spread uniformly over the 15360×8640 frame buffer, all host threads.
    keep: a list that receives (rect, RGBA8 words) of every 9th crop of the last step, for the in-run parity check."""
    O, B, kind = cpu_backend()
    envs = oracle_env_cubes(O)
    cores = O.host_threads()
    pol = O.LOD_ALWAYS_LINEAR if kind == "reference" else O.LOD_QUAD

    def run(crop):
        n, t0 = 0, time.perf_counter()
        for gy in range(9):
            for gx in range(16):
                cx = (FB_W // 16) * gx + FB_W // 32
                cy = (FB_H // 9) * gy + FB_H // 18
                x0, y0 = (cx - crop // 2) & ~1, (cy - crop // 2) & ~1
                r = B.render(SCENE_MB_POWER8, FB_W, FB_H, TIME, envs=envs, rect=(x0, y0, x0 + crop, y0 + crop), want_linear=False,
                             threads=cores, lod_policy=pol)
                if keep is not None and (gy * 16 + gx) % 9 == 4:
                    keep.append(((x0, y0, x0 + crop, y0 + crop), r["fb"]))
                n += crop * crop
        return n, time.perf_counter() - t0

    n, dt = run(64)                                   # probe, then size the sample to the budget
    rate = n / dt
    crop = int(min(960, max(64, ((rate * budget_s / 144.0) ** 0.5) // 2 * 2)))
    for _ in range(warmup):
        run(crop)
    times, n = [], 0
    for _ in range(max(1, steps)):
        if keep is not None:
            del keep[:]
        n, dt = run(crop)
        times.append(dt)
    dt = float(np.mean(times))
    return {"value": n / dt, "unit": "samples/s", "cores": cores, "kind": kind, "lod_policy": pol, "envs": envs,
            "sample": "144 crops of %dx%d samples on a 16x9 grid over the 15360x8640 frame buffer (%.2f%% of the frame); %s"
                      % (crop, crop, 100.0 * n / (FB_W * FB_H), cpu_kind_note(kind)),
            "seconds_per_step": dt, "samples_per_step": n}


def headline_config(mode_name: str, parallelism: str, l2: str) -> dict:
    """`config` of the JSON line — the same keys and workload values on both arms."""
    return {"workload": WORKLOAD, "scene": "FSMBPower8Shader", "in_time": TIME, "fb": [FB_W, FB_H], "out": [OUT_W, OUT_H],
            "spp": 1 << (2 * SPP_LOG2), "mode": mode_name, "parallelism": parallelism, "l2": l2}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # each step is a bounded sample sized so that the whole --steps/--warmup run ends within a few minutes
    w = min(args.warmup, 1)
    budget = max(0.3, min(float(os.environ.get("RMDF_BENCH_CPU_BUDGET_S", "12")), 150.0 / max(1, args.steps + w)))
    b = cpu_baseline_sample(steps=args.steps, warmup=w, budget_s=budget)
    line = {"impl": "reference", "metric": "mandelbulb_4k_samples_per_s", "value": b["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": b["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config("cpu-" + b["kind"], "%d host threads (rank 0 only)" % b["cores"], "n/a (CPU)"),
            "cpu_baseline": {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": b["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ---- our arm -------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import rmdf_b200 as R
    from importlib import import_module
    par = import_module("ray-marching-distance-fields_b200.parallel")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — librmdf_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sr = R.ShaderRenderer(local)
    cubes = build_env_cubes(sr)                              # host copies of the five cube maps
    tol = None
    if args.mode in ("auto", "fast"):
        tol = tolerance_check(sr, torch, R) if rank == 0 else None
        if world > 1:
            box = [tol]
            dist.broadcast_object_list(box, src=0)
            tol = box[0]
    use_fast = args.mode == "fast" or (args.mode == "auto" and tol["pass"])
    mode = R.MODE_FAST if use_fast else R.MODE_PARITY
    mode_name = "fast" if use_fast else "parity"
    stream = torch.cuda.Stream(device=dev)          # a real (non-NULL) stream the library and the events share
    torch.cuda.set_stream(stream)
    sr.setStream(stream.cuda_stream)

    sharded = world > 1
    scene = R.FragmentShader.FSMBPower8Shader
    full = dict(resolve_log2=SPP_LOG2, mode=mode)
    grid, skew = par.balanced_grid(FB_W, FB_H, SPP_LOG2, world) if sharded else ((0, 0), 0)
    # N > 1, three ways to assemble the frame on rank 0 (--gather):
    #   peer (default)  every rank launches the same frame-wide work list; work item i (a 32x8-sample block) belongs to rank
    #                   i mod N and goes through the rank's own queue; pixel stores straight into rank 0's frame buffer over
    #                   NVLink; arrival flags in peer memory instead of a barrier (parallel.SharedFrame.begin/end,
    #                   csrc/frame_sync.cuh)
    #   peer-queue      the same, but all GPUs claim items from ONE queue word in rank 0's memory with system-scope atomics
    #   peer-static     round 1's path: 64x64-sample tiles dealt statically, stores into rank 0's frame buffer, one NCCL barrier
    #   nccl            packed tiles + ONE dist.gather + re-assembly on rank 0
    gather = args.gather if sharded else "none"
    shared = None
    if gather in ("peer", "peer-queue", "peer-static"):
        okf = 1
        try:
            shared = par.SharedFrame(sr, OUT_W * OUT_H, world, rank)
        except R.RmdfError as e:                                   # no peer access on this box
            okf = 0
            print("bench.py: rank %d: peer frame buffer unavailable (%s)" % (rank, e), file=sys.stderr)
        okt = torch.tensor([okf], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 0:
            if shared is not None:
                shared.close()
            shared, gather = None, "nccl"
    flags_mode = gather in ("peer", "peer-queue")
    deal = 1 if gather == "peer-queue" else 0
    if flags_mode or not sharded:
        p = sr.makeParams(scene, None, FB_W, FB_H, TIME, **full)                          # the same full-frame work list on every rank
    elif gather == "peer-static":
        p = sr.makeParams(scene, None, FB_W, FB_H, TIME, shard_rank=rank, shard_count=world, layout=R.LAYOUT_FRAME, shard_grid=grid, shard_skew=skew, **full)
    else:
        p = sr.makeParams(scene, None, FB_W, FB_H, TIME, shard_rank=rank, shard_count=world, layout=R.LAYOUT_PACKED_TILES, shard_grid=grid, shard_skew=skew, **full)
    # end to end at N > 1: the shared frame in page-locked HOST memory, written by every rank's kernel directly
    host_shared = None
    if shared is not None:
        try:
            host_shared = par.SharedHostFrame(sr, OUT_W * OUT_H, world, rank)
        except (R.RmdfError, OSError) as e:
            if rank == 0:
                print("bench.py: shared host frame unavailable (%s); end to end = peer frame + D2H copy" % e, file=sys.stderr)
        okh = torch.tensor([1 if host_shared is not None else 0], device=dev)
        dist.all_reduce(okh, op=dist.ReduceOp.MIN)
        if int(okh.item()) == 0 and host_shared is not None:
            host_shared.close(); host_shared = None
    n_out, n_tiles = sr.outputPixels(p)
    out_dev = torch.empty(n_out if shared is None else 1, dtype=torch.int32, device=dev)
    shared_frame = shared.as_tensor() if (shared is not None and rank == 0) else None
    gather_list = [torch.empty(n_out, dtype=torch.int32, device=dev) for _ in range(world)] if (gather == "nccl" and rank == 0) else None
    frame_host = torch.empty(OUT_H * OUT_W, dtype=torch.int32).pin_memory() if rank == 0 else None
    cubes_pinned = [torch.from_numpy(c).pin_memory() for c in cubes]
    h2d_bytes = sum(c.numel() * 4 for c in cubes_pinned) + 64          # cube maps + rmdf_params
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def render_shared(target_ptr):
        if flags_mode:
            shared.begin(deal=deal)                                     # rank 0: queue reset + go; the others wait for go on their stream
            sr.renderDevice(p, target_ptr)                              # renders this rank's items, publishes its arrival flag
            shared.end()                                                # rank 0's stream waits for every rank's flag
        else:
            sr.renderDevice(p, target_ptr)
            shared.complete()                                           # stream sync + one NCCL barrier

    def step_resident():
        if shared is not None:
            render_shared(shared.ptr)
            return shared_frame
        sr.renderDevice(p, out_dev.data_ptr())
        if sharded:
            return par.gather_frame(out_dev, OUT_W, OUT_H, world, rank, gather_list, grid, skew)
        return out_dev

    def step_e2e(upload_assets=True):
        if upload_assets:
            sr.bindCubeMaps(range(5), cubes_pinned)                     # H2D: this step's inputs (one batch, one sync)
        if not sharded:
            sr.renderHost(p, frame_host.data_ptr())                     # kernel + D2H through the C ABI
        elif host_shared is not None:
            render_shared(host_shared.ptr)                              # every rank stores its pixels into the host frame
            torch.cuda.synchronize()                                    # rank 0: all arrival flags seen = the frame is in host memory
        else:
            frame = step_resident()
            if rank == 0:
                frame_host.copy_(frame.reshape(-1), non_blocking=True)
            torch.cuda.synchronize()

    def timed(fn, k, w, events=None):
        events = (fn is step_resident) if events is None else events
        for _ in range(w):
            flush.zero_(); fn()
        ms = []
        barrier()
        for _ in range(k):
            flush.zero_()                                               # L2 flush, outside the events
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream); fn(); e1.record(stream)
            barrier()
            ms.append(max(e0.elapsed_time(e1), 0.0) if events else (time.perf_counter() - t0) * 1e3)
        t = torch.tensor([float(np.sum(ms))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / k

    sampler = ClockSampler(local) if rank == 0 else None
    sr.resetCounters()
    l0 = sr.counters()["launches"]
    ms_res = timed(step_resident, args.steps, args.warmup)
    c = sr.counters()
    launches = (c["launches"] - l0) // (args.steps + args.warmup) * args.steps
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(step_e2e, args.steps, 2)
    # the same call sequence with the five cube maps left bound (the reference uploads them once, in withShaderRenderer, and
    # sends only uniforms per frame): per step the 64-byte parameter block goes down and the frame comes back
    ms_e2e_bound = timed(lambda: step_e2e(False), args.steps, 2, events=False)
    e2e_same = frame_same = None
    if host_shared is not None:                       # the host frame the e2e path produced == the device-resident frame
        step_resident(); step_e2e(); barrier()
        if rank == 0:
            e2e_same = bool(np.array_equal(host_shared.as_numpy().view(np.int32), shared_frame.cpu().numpy()))
    if sharded:                                       # the assembled frame == the same frame rendered by rank 0 alone
        frame = step_resident(); barrier()
        if rank == 0:
            single = torch.empty(OUT_W * OUT_H, dtype=torch.int32, device=dev)
            sr.renderDevice(sr.makeParams(scene, None, FB_W, FB_H, TIME, **full), single.data_ptr()); torch.cuda.synchronize()
            frame_same = bool(torch.equal(frame.reshape(-1), single))
            del single
        barrier()

    # BASELINE configs[4] at N > 1: the Julia frame through the same shared queue / peer stores / arrival flags
    julia_n = None
    if sharded and shared is not None and flags_mode and not args.no_configs:
        shared_j = par.SharedFrame(sr, J_OUT_W * J_OUT_H, world, rank)

        def step_julia():
            shared_j.begin(deal=deal)
            sr.juliaDevice(J_FB_W, J_FB_H, False, J_TICK, (0, J_FB_H), J_SPP_LOG2, shared_j.ptr)
            shared_j.end()
        ms_j = timed(step_julia, args.steps, 3, events=True)
        step_julia(); barrier()
        same_j = None
        if rank == 0:
            single = torch.empty(J_OUT_W * J_OUT_H, dtype=torch.int32, device=dev)
            sr.juliaDevice(J_FB_W, J_FB_H, False, J_TICK, (0, J_FB_H), J_SPP_LOG2, single.data_ptr()); torch.cuda.synchronize()
            same_j = bool(torch.equal(shared_j.as_tensor(), single))
            del single
        barrier()
        shared_j.close()
        julia_n = {"workload": J_WORKLOAD, "ms_per_step": ms_j, "value": J_FB_W * J_FB_H / (ms_j * 1e-3), "unit": "samples/s",
                   "frame_equals_single_gpu": same_j, "steps": args.steps}

    # per-launch kernel time and counters for the roofline (rank-local kernel, all ranks reduced)
    per = args.steps + args.warmup
    cnt = torch.tensor([c[k] / per for k in ("samples", "in_sphere", "hits", "march_de", "shade_de", "inner_iters")],
                       dtype=torch.float64, device=dev)
    kms = []
    for _ in range(3):
        flush.zero_(); barrier()
        step_resident(); torch.cuda.synchronize()
        kms.append(sr.counters()["last_kernel_ms"])       # this rank's march kernel alone (events around the launch)
    barrier()
    kt = torch.tensor([float(np.mean(kms))], dtype=torch.float64, device=dev)
    rank_kernel_ms = None
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        per_rank = [torch.zeros_like(kt) for _ in range(world)]
        dist.all_gather(per_rank, kt)
        rank_kernel_ms = [round(float(t_.item()), 4) for t_ in per_rank]     # how even the deal is: each rank's own kernel
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    frame_cnt = dict(zip(("samples", "in_sphere", "hits", "march_de", "shade_de", "inner_iters"), cnt.tolist()))
    kernel_ms = float(kt.item())
    # the same frame with the work order off (items claimed in scan order, as the first frame of a sequence always is):
    # every timed step above ran in the order made from the step before it (rmdf_set_work_order, frame_sync.cuh)
    sr.setWorkOrder(False)
    sms = []
    for _ in range(3):
        flush.zero_(); barrier()
        step_resident(); torch.cuda.synchronize()
        sms.append(sr.counters()["last_kernel_ms"])
    barrier()
    sr.setWorkOrder(True)
    st_ = torch.tensor([float(np.mean(sms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(st_, op=dist.ReduceOp.MAX)
    scan_order_kernel_ms = float(st_.item())
    fp32_peak = sr.measureFp32Peak() if rank == 0 else None
    other = None
    if world == 1 and use_fast:                      # the bit-exact build on the same frame, for the record
        p2 = sr.makeParams(R.FragmentShader.FSMBPower8Shader, None, FB_W, FB_H, TIME, resolve_log2=SPP_LOG2, mode=R.MODE_PARITY)
        oms = []
        for _ in range(3):
            flush.zero_(); sr.renderDevice(p2, out_dev.data_ptr()); torch.cuda.synchronize()
            oms.append(sr.counters()["last_kernel_ms"])
        other = {"mode": "parity", "kernel_ms": float(np.mean(oms[1:])), "value": FB_W * FB_H / (float(np.mean(oms[1:])) * 1e-3),
                 "unit": "samples/s", "note": "IEEE-only build, bit-identical to the CPU oracle"}

    if rank == 0:
        peaks, peaks_src = load_peaks()
        flops = algorithmic_flops(frame_cnt)
        samples = FB_W * FB_H
        ach_tf = flops / (kernel_ms * 1e-3) / 1e12 / world               # per-GPU rate
        traffic = None
        try:                                                             # from the committed ncu capture (per launch, N = 1)
            tr = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())[mode_name]
            traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) if world == 1 else None
        except Exception:
            pass
        alg_bytes = OUT_W * OUT_H * 4 + sum(6 * c_.shape[1] ** 2 * 8 for c_ in cubes)    # resolved frame + half4 texels
        line = {
            "metric": "mandelbulb_4k_samples_per_s", "value": samples / (ms_res * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config(mode_name, {"none": "single GPU",
                                                  "peer": "one frame-wide work list, item i (32x8 samples) -> rank i mod %d through the rank's own queue; pixel stores into rank 0's frame buffer over NVLink; arrival flags in peer memory, no barrier, no NCCL on the data path" % world,
                                                  "peer-queue": "one frame-wide work queue in rank 0's memory, claimed by all %d GPUs with system-scope atomics over NVLink (dynamic deal); pixel stores into rank 0's frame buffer; arrival flags, no barrier" % world,
                                                  "peer-static": "tiles %dx%d, rank = (tx + %d*ty) mod %d, stores into rank 0's frame buffer over NVLink + one NCCL barrier" % (grid[0], grid[1], skew, world),
                                                  "nccl": "tiles %dx%d, rank = (tx + %d*ty) mod %d, one NCCL gather + re-assembly" % (grid[0], grid[1], skew, world)}[gather],
                                      "flushed between steps (256 MiB write); env maps (2.8 MB) are meant to be L2-resident"),
            "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": OUT_W * OUT_H * 4,
                    "path": ("rmdf_render with a page-locked host frame buffer: the kernel's stores are the D2H" if not sharded else
                             ("every rank's kernel stores its pixels into one page-locked host frame (POSIX shm) over its own PCIe link; frame equals the device-resident one: %s" % e2e_same) if host_shared is not None
                             else "device-resident gather + D2H copy on rank 0")},
            "e2e_assets_bound": {"value": samples / (ms_e2e_bound * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e_bound,
                                 "h2d_bytes_per_step": 64, "d2h_bytes_per_step": OUT_W * OUT_H * 4,
                                 "note": "as e2e, but the cube maps stay bound between steps (uploaded once, as the reference's withShaderRenderer does); "
                                         "e2e above re-uploads all five on every rank in every step"},
            "gpu_launches": int(launches),
            "sdf_steps_per_s": (frame_cnt["march_de"] + frame_cnt["shade_de"]) / (ms_res * 1e-3),
            "frame_counters": {k: int(v) for k, v in frame_cnt.items()},
            "roofline": {"bound": "fp32_fma", "achieved": ach_tf, "peak": fp32_peak["tflops"], "unit": "TFLOP/s",
                         "frac": ach_tf / fp32_peak["tflops"], "traffic": traffic,
                         "peak_source": "in-run FFMA-chain microbenchmark (148 SM x 128 lanes x 2 x clock); MEASURED_PEAKS.json carries no FP32 figure",
                         "algorithmic_flop_per_frame": flops, "kernel_ms": kernel_ms,
                         "hbm": {"achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peaks_src}},
            "work_order": {"enabled": True, "kernel_ms": kernel_ms, "scan_order_kernel_ms": scan_order_kernel_ms,
                           "note": "every step claims its work items most-expensive-first, in the order a small kernel made from the previous "
                                   "step's measured item costs (scheduling only: same pixels, same counters); kernel_ms includes that kernel; "
                                   "scan_order_kernel_ms = the same frame with the order switched off, as a first frame is rendered"},
            "clocks": clocks,
        }
        if sharded:
            line["frame_equals_single_gpu"] = frame_same
            line["rank_kernel_ms"] = rank_kernel_ms
            if julia_n is not None:
                line["julia_8k_64spp"] = julia_n
        if tol is not None:
            line["tolerance_check"] = tol
        if other is not None:
            line["other_mode"] = other
        if world == 1 and not args.no_cpu_baseline:
            kept = []
            b = cpu_baseline_sample(keep=kept)
            line["cpu_baseline"] = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
            # parity in this very run, at the real frame size: the CPU leg's pixels (the reference's shader compiled by
            # g++ when kind = "reference") against the CUDA parity build's unresolved frame, same filtering policy
            # (the CPU leg's own cube maps are bound for this: the GPU-built ones differ from them in the last bits of some
            # texels — convolution summation order, CUDA vs glibc acosf — which is a property of the prefilter)
            pp = sr.makeParams(scene, None, FB_W, FB_H, TIME, mode=R.MODE_PARITY, lod_policy=b["lod_policy"])
            lvl0 = torch.empty(FB_W * FB_H, dtype=torch.int32, device=dev)
            sr.bindCubeMaps(range(5), b["envs"])
            sr.renderDevice(pp, lvl0.data_ptr()); torch.cuda.synchronize()
            sr.bindCubeMaps(range(5), cubes)
            frame0 = lvl0.reshape(FB_H, FB_W)
            bad = tot = 0
            for (x0, y0, x1, y1), want in kept:
                got = frame0[y0:y1, x0:x1].cpu().numpy().view(np.uint32)
                bad += int((got != want).sum()); tot += want.size
            del lvl0, frame0
            line["parity_check"] = {"against": "cpu_baseline kind '%s' (reference = /root/reference/fragment.shd compiled by g++, oracle/_ref)" % b["kind"],
                                    "what": "RGBA8 words of the CUDA parity build, unresolved 15360x8640 frame buffer, %d crops spread over the frame" % len(kept),
                                    "samples_compared": tot, "samples_differing": bad, "bit_identical": bool(tot > 0 and bad == 0)}
        if world == 1 and not args.no_configs:
            # the other BASELINE configs (1, 2, 3, 5) and the fourth scene, a few hundred ms of GPU time in all
            import bench_configs
            line["configs"] = bench_configs.measure(sr, cubes, torch, R, flush, local, stream, with_cpu=not args.no_cpu_baseline)
        emit(line)
    if host_shared is not None:
        host_shared.close()
    if shared is not None:
        torch.cuda.synchronize(); dist.barrier(); shared.close()
    sr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---- secondary workload: BASELINE.json configs[4], the 2D Julia set ---------------------------
J_OUT_W, J_OUT_H, J_SPP_LOG2, J_TICK = 7680, 4320, 3, 7.0          # 64 spp = 8×8 ordered grid
J_FB_W, J_FB_H = J_OUT_W << J_SPP_LOG2, J_OUT_H << J_SPP_LOG2      # 61440 × 34560 samples
J_WORKLOAD = "Julia set 7680x4320 x64spp (FB 61440x34560), juliaAnimated tick 7, plain colouring"
F_JULIA_ITER = 13.0                                                # flop per iteration, SURVEY.md §8d


def julia_cpu_baseline(steps: int = 1, budget_s: float = float(os.environ.get("RMDF_BENCH_CPU_BUDGET_S", "12"))):
    """The CPU restatement of Fractal2D.hs over 135 row slabs spread uniformly over the 34560 rows, each slab
    split the reference's way: makeNSegments into one contiguous row segment per host thread
    (forSegmentsConcurrently Nothing, Fractal2D.hs:98 / ConcurrentSegments.hs:14-28)."""
    from oracle import oracle as O
    cores = O.host_threads()
    unit = max(32, cores)                              # at least one row per segment

    def run(rows_per_slab):
        n, t0 = 0, time.perf_counter()
        for i in range(135):
            y0 = i * 256
            O.julia(J_FB_W, J_FB_H, False, J_TICK, rows=(y0, y0 + rows_per_slab), segments=cores)
            n += rows_per_slab * J_FB_W
        return n, time.perf_counter() - t0
    n, dt = run(unit)                                  # probe
    rows = int(max(unit, min(256, (n / dt) * budget_s / (135 * J_FB_W))))
    times = []
    for _ in range(max(1, steps)):
        n, dt = run(rows); times.append(dt)
    dt = float(np.mean(times))
    return {"value": n / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "135 slabs of %d rows x 61440 samples spread over the frame (%.2f%% of it), each cut by makeNSegments into %d row segments, one thread each "
                      "(the reference's forSegmentsConcurrently); C restatement of Fractal2D.hs — not GHC" % (rows, 100.0 * n / (J_FB_W * J_FB_H), cores),
            "seconds_per_step": dt}


def run_julia(args):
    import torch
    import torch.distributed as dist
    import rmdf_b200 as R
    from importlib import import_module
    par = import_module("ray-marching-distance-fields_b200.parallel")
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            b = julia_cpu_baseline(steps=args.steps)
            emit({"impl": "reference", "metric": "julia_8k_samples_per_s", "value": b["value"], "unit": "samples/s", "n_gpus": args.gpus,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": b["seconds_per_step"] * 1e3, "higher_is_better": True,
                              "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": J_WORKLOAD},
                              "cpu_baseline": {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")},
                              "e2e": {"value": b["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — librmdf_b200 has no CPU fallback")
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sr = R.ShaderRenderer(local)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); sr.setStream(stream.cuda_stream)
    band_h = par.julia_band_height(J_FB_H, J_SPP_LOG2, world)
    n_out = J_OUT_W * J_OUT_H // world
    shared = None
    if world > 1 and args.gather in ("peer", "peer-queue", "peer-static"):     # fused gather, as for the Mandelbulb frame
        okf = 1
        try:
            shared = par.SharedFrame(sr, J_OUT_W * J_OUT_H, world, rank)
        except R.RmdfError:
            okf = 0
        okt = torch.tensor([okf], device=dev); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 0 and shared is not None:
            shared.close(); shared = None
    shared_frame = shared.as_tensor() if (shared is not None and rank == 0) else None
    out_dev = torch.empty(n_out if shared is None else 1, dtype=torch.int32, device=dev)
    gather_list = [torch.empty(n_out, dtype=torch.int32, device=dev) for _ in range(world)] if (world > 1 and rank == 0 and shared is None) else None
    frame_host = torch.empty(J_OUT_H * J_OUT_W, dtype=torch.int32).pin_memory() if rank == 0 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if shared is not None and args.gather in ("peer", "peer-queue"):     # one frame-wide item list for all GPUs (parallel.SharedFrame)
            shared.begin(deal=1 if args.gather == "peer-queue" else 0)
            sr.juliaDevice(J_FB_W, J_FB_H, False, J_TICK, (0, J_FB_H), J_SPP_LOG2, shared.ptr)
            shared.end()
            return shared_frame
        if shared is not None:                                          # static interleaved bands + barrier
            sr.juliaBandsDevice(J_FB_W, J_FB_H, False, J_TICK, J_SPP_LOG2, band_h, rank, world, shared.ptr, layout=R.LAYOUT_FRAME)
            shared.complete()
            return shared_frame
        sr.juliaBandsDevice(J_FB_W, J_FB_H, False, J_TICK, J_SPP_LOG2, band_h, rank, world, out_dev.data_ptr())
        return par.gather_bands(out_dev, J_OUT_W, J_OUT_H, world, rank, band_h >> J_SPP_LOG2, gather_list)

    def step_e2e():
        if world == 1:                                                  # the reference-facing call: host buffer in, frame out
            import ctypes as C
            R._capi.check(sr._lib.rmdf_julia(sr.handle, J_FB_W, J_FB_H, 0, C.c_double(J_TICK), 0, J_FB_H, J_SPP_LOG2,
                                             C.c_void_p(frame_host.data_ptr())))
            return
        frame = step_resident()
        if rank == 0:
            frame_host.copy_(frame.reshape(-1), non_blocking=True)
        torch.cuda.synchronize()

    def timed(fn, k, w, events):
        for _ in range(w):
            flush.zero_(); fn()
        ms = []
        for _ in range(k):
            flush.zero_(); barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record(stream); fn(); e1.record(stream); barrier()
            ms.append(e0.elapsed_time(e1) if events else (time.perf_counter() - t0) * 1e3)
        t = torch.tensor([float(np.sum(ms))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / k

    sampler = ClockSampler(local) if rank == 0 else None
    sr.resetCounters()
    ms_res = timed(step_resident, args.steps, args.warmup, True)
    c = sr.counters()
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(step_e2e, args.steps, 2, False)
    its = torch.tensor([c["inner_iters"] / (args.steps + args.warmup)], dtype=torch.float64, device=dev)
    kt = torch.tensor([c["last_kernel_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(its, op=dist.ReduceOp.SUM); dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak = sr.measureFp32Peak()["tflops"]
        samples = J_FB_W * J_FB_H
        ach = float(its.item()) * F_JULIA_ITER / (float(kt.item()) * 1e-3) / 1e12 / world
        line = {"metric": "julia_8k_samples_per_s", "value": samples / (ms_res * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": J_WORKLOAD, "fb": [J_FB_W, J_FB_H], "out": [J_OUT_W, J_OUT_H], "spp": 64,
                                                "parallelism": ("one frame-wide item list, item i -> rank i mod %d (peer-queue: one shared queue), stores into rank 0's frame over NVLink, arrival flags" % world if (shared is not None and args.gather in ("peer", "peer-queue")) else
                                                                "row bands of %d rows, band b -> rank b mod %d, %s" % (band_h, world, "stores fused into the kernel over NVLink peer memory + barrier" if shared is not None else "one NCCL gather")) if world > 1 else "single GPU",
                                                "l2": "flushed between steps (256 MiB write)"},
                "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 64,
                        "d2h_bytes_per_step": J_OUT_W * J_OUT_H * 4},
                "gpu_launches": args.steps, "iterations_per_frame": int(its.item()),
                "roofline": {"bound": "fp32_fma", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                             "kernel_ms": float(kt.item()), "note": "13 flop per z <- z^2 + c iteration; 3.8 iterations per sample on this frame, so per-sample bookkeeping weighs as much as the orbit (ncu: profiles/r1_julia_v3_packed.md)"},
                "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            b = julia_cpu_baseline()
            line["cpu_baseline"] = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    if shared is not None:
        torch.cuda.synchronize(); dist.barrier(); shared.close()
    sr.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "parity", "fast"],
                    help="auto: the fast build if it meets BASELINE's tolerance against the parity build in this run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary BASELINE configs (N = 1 only)")
    ap.add_argument("--gather", default="peer", choices=["peer", "peer-queue", "peer-static", "nccl"],
                    help="N > 1: interleaved items + stores over NVLink peer memory + arrival flags (default); the same with one shared "
                         "work queue; round 1's static tiles + barrier; NCCL gather")
    ap.add_argument("--workload", default="mandelbulb4k", choices=["mandelbulb4k", "julia8k"],
                    help="mandelbulb4k = BASELINE's headline metric (default); julia8k = BASELINE configs[4]")
    args = ap.parse_args()
    _claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "julia8k" and (args.impl == "reference" or int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.gpus == 1):
        return run_julia(args)
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__] + sys.argv[1:]
        return subprocess.call(cmd, stdout=_RESULT_OUT)
    return run_julia(args) if args.workload == "julia8k" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
