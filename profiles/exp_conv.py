"""convolve sweep timing per library variant: python profiles/exp_conv.py NAME..."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if os.environ.get("RMDF_EXP_CHILD"):
    sys.path.insert(0, str(ROOT))
    import numpy as np, torch
    import rmdf_b200 as R
    with R.ShaderRenderer(0) as sr:
        refl = R.loadHDRImage(ROOT / "data" / "uffizi_512.hdr")
        small = sr.resizeHDRImage(refl, 256)
        d = torch.from_numpy(small).cuda(); o = torch.empty((4,) + small.shape, device="cuda")
        pw = (R._capi.C.c_float * 4)(1, 8, 64, 512)
        ms = []
        for i in range(10):
            R._capi.check(sr._lib.rmdf_envmap_convolve_device(sr.handle, d.data_ptr(), 256, 128, pw, 4, o.data_ptr())); sr.sync()
            ms.append(sr.counters()["last_kernel_ms"])
        print(os.environ["RMDF_VARIANT"], "convolve sweep ms: median %.4f best %.4f" % (float(np.median(ms[2:])), min(ms[2:])), flush=True)
else:
    for name in sys.argv[1:]:
        lib = ROOT / "ray-marching-distance-fields_b200" / ("librmdf_b200.so" if name == "default" else "librmdf_b200_%s.so" % name)
        subprocess.run([sys.executable, __file__], env=dict(os.environ, RMDF_LIB=str(lib), RMDF_VARIANT=name, RMDF_EXP_CHILD="1"))
