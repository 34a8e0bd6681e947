import sys; sys.path.insert(0, ".")
import torch, rmdf_b200 as R
smooth = int(sys.argv[1]) if len(sys.argv) > 1 else 0
with R.ShaderRenderer(0) as sr:
    w, h, k = 61440, 34560, 3
    out = torch.empty((w >> k) * (h >> k), dtype=torch.int32, device="cuda")
    for i in range(3):
        sr.juliaDevice(w, h, smooth, 7.0, (0, h), k, out.data_ptr()); sr.sync()
    print("kernel ms", sr.counters()["last_kernel_ms"])
