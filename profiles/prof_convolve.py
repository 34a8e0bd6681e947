"""Config 2: the prefilter sweep (4 powers of the 256x128 downsample) through rmdf_envmap_convolve_device."""
import sys
sys.path.insert(0, ".")
import torch
import rmdf_b200 as R

with R.ShaderRenderer(0) as sr:
    refl = R.loadHDRImage("data/uffizi_512.hdr")
    small = sr.resizeHDRImage(refl, 256)
    dsmall = torch.from_numpy(small).cuda(); dout = torch.empty((4,) + small.shape, device="cuda")
    pw = (R._capi.C.c_float * 4)(1, 8, 64, 512)
    ms = []
    for i in range(8):
        R._capi.check(sr._lib.rmdf_envmap_convolve_device(sr.handle, dsmall.data_ptr(), 256, 128, pw, 4, dout.data_ptr())); sr.sync()
        ms.append(sr.counters()["last_kernel_ms"])
    print("convolve sweep ms: first %.3f, best %.3f" % (ms[0], min(ms[1:])))
