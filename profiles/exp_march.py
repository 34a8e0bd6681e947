"""Quick kernel-time probe on the GPU box: C4 frame, both modes; prints kernel ms (library
CUDA events) and, with --psnr, fast-vs-parity agreement on a level-0 crop."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import rmdf_b200 as R
from bench import build_env_cubes, FB_W, FB_H

scenes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2]
with R.ShaderRenderer(0) as sr:
    build_env_cubes(sr)
    import torch
    for scene in scenes:
        for mode in (R.MODE_PARITY, R.MODE_FAST):
            p = sr.makeParams(scene, None, FB_W, FB_H, 0.0, resolve_log2=2, mode=mode)
            n, _ = sr.outputPixels(p)
            out = torch.empty(n, dtype=torch.int32, device="cuda")
            ms = []
            for i in range(4):
                sr.renderDevice(p, out.data_ptr()); sr.sync()
                ms.append(sr.counters()["last_kernel_ms"])
            print("scene %d mode %s: kernel ms %s  -> %.3f Gsamples/s" % (scene, "fast" if mode else "parity", ["%.2f" % m for m in ms], FB_W * FB_H / min(ms[1:]) / 1e6))
    if "--psnr" in sys.argv:
        # fast vs parity (= oracle, bit-exact) on the full frame at 1/4 scale and on the C4 frame's centre crop
        for (w, h, k) in [(3840, 2160, 0), (7680, 4320, 1), (15360, 8640, 2)]:
            # render a horizontal band through the centre to bound memory: use tiles 24..39 (rows 3,4)
            accs = {}
            for mode in (R.MODE_PARITY, R.MODE_FAST):
                bands = []
                for idx in range(24, 40):
                    p = sr.makeParams(2, idx, w, h, 0.0, out_format=R.OUT_RGBA32F, mode=mode, layout=R.LAYOUT_PACKED_TILES)
                    n, _ = sr.outputPixels(p)
                    out = torch.empty(n * 4, dtype=torch.float32, device="cuda")
                    sr.renderDevice(p, out.data_ptr()); sr.sync()
                    t = out.reshape(h // 8, w // 8, 4)
                    s = 1 << k
                    lin = torch.nan_to_num(t[..., :3]).clamp(0, 1)
                    lin = lin.reshape(h // 8 // s, s, w // 8 // s, s, 3).mean(dim=(1, 3))
                    hit = t[..., 3]
                    bands.append((lin.cpu().numpy(), hit.cpu().numpy()))
                accs[mode] = bands
            mse = np.mean([np.mean((a[0].astype(np.float64) - b[0]) ** 2) for a, b in zip(accs[0], accs[1])])
            agree = np.mean([np.mean(a[1] == b[1]) for a, b in zip(accs[0], accs[1])])
            print("fast vs parity, FB %dx%d resolved /%d (tiles 24-39): hit agreement %.5f%%, linear PSNR (clamped [0,1]) %.2f dB" % (w, h, 1 << k, 100 * agree, 10 * np.log10(1 / mse)))
