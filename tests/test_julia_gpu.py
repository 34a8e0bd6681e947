"""GPU parity of the Julia kernel (csrc/julia.cu) against the oracle's restatement of
Fractal2D.hs:63-98.  Integer output: bit-exact for the plain variant; the smooth variant goes
through log(log|z|²) (libm in the reference, deterministic rg_log here) and may differ by one
8-bit step on a vanishing fraction of pixels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tick", [0.0, 7.0, 123.456])
@pytest.mark.parametrize("size", [(256, 144), (97, 33), (1, 1)])
def test_julia_bit_exact(sr, O, tick, size):
    w, h = size
    assert np.array_equal(sr.juliaAnimated(w, h, None, False, tick), O.julia(w, h, False, tick))


@pytest.mark.parametrize("tick", [0.0, 7.0])
def test_julia_smooth(sr, O, tick):
    w, h = 512, 288
    got, ref = sr.juliaAnimated(w, h, None, True, tick), O.julia(w, h, True, tick)
    d = np.abs((got >> 8).astype(np.int64) - (ref >> 8).astype(np.int64))
    print("julia smooth tick %g: %.5f%% pixels differ, max step %d" % (tick, 100 * (d > 0).mean(), d.max()))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    assert ((got & 0xFFFF00FF) == 0).all()


def test_julia_rows_and_resolve(sr, O):
    """BASELINE config 5 shape in miniature: 64 spp (8×8) resolved in-kernel == mip level 3 of level 0;
    a row band (multi-GPU decomposition by rows, like forSegmentsConcurrently) == the same rows."""
    w, h = 512, 256
    full = O.julia(w, h, False, 7.0)
    for k in (1, 2, 3):
        assert np.array_equal(sr.juliaAnimated(w, h, None, False, 7.0, resolve_log2=k), O.mip_resolve(full, k)), k
    band = sr.juliaAnimated(w, h, None, False, 7.0, rows=(64, 192))
    assert np.array_equal(band, full[64:192])
    band2 = sr.juliaAnimated(w, h, None, False, 7.0, rows=(64, 192), resolve_log2=3)
    assert np.array_equal(band2, O.mip_resolve(full, 3)[8:24])
    c = sr.counters()
    assert c["samples"] > 0 and c["inner_iters"] > 0


@pytest.mark.parametrize("count", [1, 2, 4, 8])
def test_julia_bands_partition_the_frame(sr, O, count):
    """Multi-GPU decomposition on one GPU: interleaved row bands of all ranks reassemble the frame."""
    import torch
    from importlib import import_module
    par = import_module("ray-marching-distance-fields_b200.parallel")
    w, h, k = 512, 1024, 2
    band_h = par.julia_band_height(h, k, count, min_bands_per_rank=2)
    assert h % band_h == 0 and (h // band_h) % count == 0
    want = O.mip_resolve(O.julia(w, h, False, 7.0), k)
    packed = []
    for r in range(count):
        out = torch.zeros((h >> k) // count * (w >> k), dtype=torch.int32, device="cuda")
        sr.juliaBandsDevice(w, h, False, 7.0, k, band_h, r, count, out.data_ptr()); sr.sync()
        packed.append(out.cpu().numpy().view(np.uint32))
    got = par.assemble_bands_numpy(packed, w >> k, h >> k, count, band_h >> k)
    assert np.array_equal(got, want)
    # frame layout: every rank writes its bands in place into one shared-size buffer
    frame = torch.zeros((h >> k) * (w >> k), dtype=torch.int32, device="cuda")
    for r in range(count):
        sr.juliaBandsDevice(w, h, False, 7.0, k, band_h, r, count, frame.data_ptr(), layout=0)      # RMDF_LAYOUT_FRAME
    sr.sync()
    assert np.array_equal(frame.cpu().numpy().view(np.uint32).reshape(h >> k, w >> k), want)


def test_julia_errors(sr, R):
    with pytest.raises(R.RmdfError):
        sr.juliaAnimated(100, 100, None, False, 0.0, resolve_log2=3)        # 100 % 8 != 0
    with pytest.raises(R.RmdfError):
        sr.juliaAnimated(64, 64, None, False, 0.0, rows=(32, 16))
