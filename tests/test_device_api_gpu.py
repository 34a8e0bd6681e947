"""Device-pointer entry points of the C ABI (what a host that keeps frames on the GPU calls):
rmdf_render_device + rmdf_set_stream + rmdf_mip_resolve_device + rmdf_envmap_convolve_device,
driven with torch tensors as plain device memory."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_render_device_on_caller_stream_and_mip_chain(sr, O, R, oracle_envs):
    import torch
    for s in range(5):
        sr.bindCubeMap(s, oracle_envs[s])
    w, h = 512, 256
    stream = torch.cuda.Stream()
    sr.setStream(stream.cuda_stream)
    try:
        with torch.cuda.stream(stream):
            lvl0 = torch.zeros(w * h, dtype=torch.int32, device="cuda")
            p = sr.makeParams(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sr.renderDevice(p, lvl0.data_ptr())                    # asynchronous on the caller's stream
            e1.record(stream)
            outs = {}
            for k in (1, 2, 3):
                outs[k] = torch.zeros((w >> k) * (h >> k), dtype=torch.int32, device="cuda")
                R._capi.check(sr._lib.rmdf_mip_resolve_device(sr.handle, lvl0.data_ptr(), w, h, k, outs[k].data_ptr()))
        stream.synchronize()
        assert e0.elapsed_time(e1) > 0.0                           # the kernel really ran on that stream
        ref = O.render(2, w, h, 0.0, envs=oracle_envs)["fb"]
        assert np.array_equal(lvl0.cpu().numpy().view(np.uint32).reshape(h, w), ref)
        for k in (1, 2, 3):
            got = outs[k].cpu().numpy().view(np.uint32).reshape(h >> k, w >> k)
            assert np.array_equal(got, O.mip_resolve(ref, k)), k
            # … and the standalone chain equals the in-kernel resolve
            assert np.array_equal(got, sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k))
    finally:
        sr.setStream(None)
    with pytest.raises(R.RmdfError):
        R._capi.check(sr._lib.rmdf_mip_resolve_device(sr.handle, None, w, h, 1, None))


def test_convolve_device_pointers(sr, O, R, uffizi):
    import torch
    small = O.resize(uffizi, 64)
    src = torch.from_numpy(small).cuda()
    out = torch.zeros((2,) + small.shape, dtype=torch.float32, device="cuda")
    pw = (C.c_float * 2)(1.0, 8.0)
    R._capi.check(sr._lib.rmdf_envmap_convolve_device(sr.handle, src.data_ptr(), small.shape[1], small.shape[0], pw, 2, out.data_ptr()))
    sr.sync()
    got = out.cpu().numpy()
    assert np.allclose(got[0], O.convolve(small, 1.0), rtol=1e-4) and np.allclose(got[1], O.convolve(small, 8.0), rtol=1e-4)
    assert sr.counters()["last_kernel_ms"] > 0.0


def test_convolve_row_ranges_compose(sr, O, R, uffizi):
    """the sweep shards by destination rows; every rank reads the whole source.  Row
    ranges computed separately equal the one-pass result bit for bit (same per-pixel summation order)."""
    import torch
    small = O.resize(uffizi, 96)                                  # 96×48
    h, w = small.shape[:2]
    src = torch.from_numpy(small).cuda()
    pw = (C.c_float * 4)(1.0, 8.0, 64.0, 512.0)
    whole = torch.zeros((4, h, w, 3), dtype=torch.float32, device="cuda")
    R._capi.check(sr._lib.rmdf_envmap_convolve_device(sr.handle, src.data_ptr(), w, h, pw, 4, whole.data_ptr()))
    parts = torch.zeros_like(whole)
    for y0, y1 in ((0, 7), (7, 24), (24, 48)):                    # ragged against the 8-row CTA tile
        R._capi.check(sr._lib.rmdf_envmap_convolve_rows_device(sr.handle, src.data_ptr(), w, h, pw, 4, y0, y1, parts.data_ptr()))
    sr.sync()
    assert torch.equal(parts, whole)
    with pytest.raises(R.RmdfError):
        R._capi.check(sr._lib.rmdf_envmap_convolve_rows_device(sr.handle, src.data_ptr(), w, h, pw, 4, 10, 10, parts.data_ptr()))


def test_counters_accumulate_and_reset(sr, R):
    sr.resetCounters()
    sr.drawShaderTile(R.FragmentShader.FSDETestShader, None, 64, 64, 0.0)
    a = sr.counters()
    sr.drawShaderTile(R.FragmentShader.FSDETestShader, None, 64, 64, 0.0)
    b = sr.counters()
    assert a["samples"] == 64 * 64 and b["samples"] == 2 * 64 * 64 and b["march_de"] == 2 * a["march_de"]
    assert b["launches"] == a["launches"] + 1
    sr.resetCounters()
    assert sr.counters()["samples"] == 0
    pk = sr.measureFp32Peak()
    assert 30.0 < pk["tflops"] < 90.0                              # 148 SM × 128 lanes × 2 × ~1.9 GHz ≈ 72


def test_page_locked_host_buffers_are_written_in_place(sr, O, R, oracle_envs):
    """rmdf_render with a page-locked host buffer: the kernel stores straight into it (no staging copy); the
    pixels are the same as through pageable memory and a tile leaves the rest of the buffer alone.  rmdf_julia
    accepts page-locked memory too (through its staging buffer)."""
    import torch
    for s in range(5):
        sr.bindCubeMap(s, oracle_envs[s])
    w, h, k = 512, 256, 1
    want = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)      # pageable numpy buffer
    pinned = torch.full(((h >> k) * (w >> k),), 0x55, dtype=torch.int32).pin_memory()
    p = sr.makeParams(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
    sr.renderHost(p, pinned.data_ptr())
    assert np.array_equal(pinned.numpy().view(np.uint32).reshape(h >> k, w >> k), want)
    # one tile of the reference's 8×8 grid into a pre-filled pinned frame: everything else keeps the caller's value
    pinned.fill_(0x55)
    pt = sr.makeParams(R.FragmentShader.FSMBPower8Shader, 27, w, h, 0.0, resolve_log2=k)
    sr.renderHost(pt, pinned.data_ptr())
    got = pinned.numpy().view(np.uint32).reshape(h >> k, w >> k)
    tw, th = (w >> k) // 8, (h >> k) // 8
    tx, ty = 27 % 8, 27 // 8
    mask = np.zeros_like(got, bool); mask[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw] = True
    assert np.array_equal(got[mask], want[mask]) and (got[~mask] == 0x55).all()
    # Julia
    jw, jh = 256, 144
    jp = torch.zeros(jw * jh, dtype=torch.int32).pin_memory()
    R._capi.check(sr._lib.rmdf_julia(sr.handle, jw, jh, 0, C.c_double(7.0), 0, jh, 0, C.c_void_p(jp.data_ptr())))
    assert np.array_equal(jp.numpy().view(np.uint32).reshape(jh, jw), O.julia(jw, jh, False, 7.0))


def test_host_registered_frame_buffer(sr, R, oracle_envs):
    """rmdf_host_register: ordinary host memory (here a numpy array; in the multi-rank bench a POSIX shared-memory
    segment) page-locked for the device and used as the render target of rmdf_render_device — what every rank
    does with the shared host frame (parallel.SharedHostFrame)."""
    for s in range(5):
        sr.bindCubeMap(s, oracle_envs[s])
    w, h, k = 512, 256, 1
    want = sr.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
    buf = np.zeros((h >> k, w >> k), np.uint32)
    dev_ptr = sr.hostRegister(buf.ctypes.data, buf.nbytes)
    try:
        for rank in range(2):                                   # two "ranks" fill their tiles of the same host frame
            p = sr.makeParams(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k, shard_rank=rank, shard_count=2,
                              shard_grid=(16, 8), shard_skew=1)
            sr.renderDevice(p, dev_ptr)
        sr.sync()
    finally:
        sr.hostUnregister(buf.ctypes.data)
    assert np.array_equal(buf, want)


def _view_i32(ptr, n):
    import torch

    class _W:
        pass
    w = _W()
    w.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}
    return torch.as_tensor(w, device="cuda:0")


@pytest.mark.parametrize("deal", [0, 1])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_shared_queue_frame_protocol_on_one_gpu(R, O, oracle_envs, world, deal):
    """include/rmdf.h rmdf_frame_begin / rmdf_frame_end (csrc/frame_sync.cuh): `world` contexts on ONE device play
    the ranks of a multi-GPU frame — one control block, one frame buffer, every context launches the SAME full-frame
    work list, the kernels (concurrent, on their own streams) claim items from the one queue, the last CTA of each
    publishes the frame number, rank 0's stream waits for all arrival flags.  The frame must be bit-identical to a
    plain render, every sample rendered exactly once, for consecutive frames and for the Julia kernel too.
    deal 0 = interleaved (item i → rank i mod world, local queues), 1 = the one shared queue."""
    import torch
    w, h, k = 1024, 512, 1
    n_out = (w >> k) * (h >> k)
    srs = [R.ShaderRenderer(0) for _ in range(world)]
    try:
        for r in srs:
            for s in range(5):
                r.bindCubeMap(s, oracle_envs[s])
        want = srs[0].drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
        frame_bytes = (n_out * 4 + 255) & ~255
        ptr, _ = srs[0].sharedAlloc(frame_bytes + 64 * 4)
        ctrl = ptr + frame_bytes
        p = srs[0].makeParams(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
        for seq in (1, 2, 3):
            _view_i32(ptr, n_out).zero_(); torch.cuda.synchronize()
            for r in srs:
                r.resetCounters()
            for rank in reversed(range(world)):               # the waiting ranks first: rank 0's "go" must still get through
                srs[rank].frameBegin(ctrl, rank, world, seq, 5000, deal)
                srs[rank].renderDevice(p, ptr)
            srs[0].frameEnd(ctrl, world, seq, 5000)
            assert srs[0].frameSyncStatus() == 0              # synchronises rank 0's stream: the frame is complete
            got = _view_i32(ptr, n_out).cpu().numpy().view(np.uint32).reshape(h >> k, w >> k)
            assert np.array_equal(got, want), seq
            for r in srs[1:]:
                assert r.frameSyncStatus() == 0
            cs = [r.counters()["samples"] for r in srs]
            assert sum(cs) == w * h                                            # every sample exactly once
            if deal == 0:
                assert max(cs) - min(cs) <= 32 * 8 * 16                        # interleaved: equal shares to within a group of items or two
        # Julia through the same protocol
        jw, jh = w, h                                         # the same output size: the shared buffer holds n_out words
        wantj = srs[0].juliaAnimated(jw, jh, None, False, 7.0, resolve_log2=k)
        for rank in reversed(range(world)):
            srs[rank].frameBegin(ctrl, rank, world, 4, 5000, deal)
            srs[rank].juliaDevice(jw, jh, False, 7.0, (0, jh), k, ptr)
        srs[0].frameEnd(ctrl, world, 4, 5000)
        assert srs[0].frameSyncStatus() == 0
        gotj = _view_i32(ptr, (jw >> k) * (jh >> k)).cpu().numpy().view(np.uint32).reshape(jh >> k, jw >> k)
        assert np.array_equal(gotj, wantj)
        # a tile / shard / packed layout cannot be combined with the shared queue
        srs[0].frameBegin(ctrl, 0, 1, 5, 1000)
        with pytest.raises(R.RmdfError):
            srs[0].renderDevice(srs[0].makeParams(R.FragmentShader.FSMBPower8Shader, 3, w, h, 0.0), ptr)
        srs[0].renderDevice(p, ptr); srs[0].frameEnd(ctrl, 1, 5, 1000)
        assert srs[0].frameSyncStatus() == 0
        # a wait that cannot be satisfied gives up after its timeout and says which word it waited for
        srs[0].frameEnd(ctrl, 2, 99, 50)
        assert srs[0].frameSyncStatus() in (1, 2)
        for r in srs:
            r.sync()
        srs[0].sharedFree(ptr)
    finally:
        for r in srs:
            r.close()


@pytest.mark.parametrize("deal", [0, 1])
def test_shared_queue_frame_two_gpus(R, oracle_envs, deal):
    """The same protocol across two devices over peer memory (single process; run with gpurun --gpus 2)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w, h, k = 8192, 4096, 2                                   # large enough that both GPUs get work before the queue is empty
    n_out = (w >> k) * (h >> k)
    with R.ShaderRenderer(0) as a, R.ShaderRenderer(1) as b:
        for r in (a, b):
            for s in range(5):
                r.bindCubeMap(s, oracle_envs[s])
        b.enablePeer(0)
        want = a.drawShaderTile(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
        frame_bytes = (n_out * 4 + 255) & ~255
        ptr, _ = a.sharedAlloc(frame_bytes + 64 * 4)
        ctrl = ptr + frame_bytes
        p = a.makeParams(R.FragmentShader.FSMBPower8Shader, None, w, h, 0.0, resolve_log2=k)
        for seq in (1, 2):
            a.resetCounters(); b.resetCounters()
            b.frameBegin(ctrl, 1, 2, seq, 5000, deal); b.renderDevice(p, ptr)
            a.frameBegin(ctrl, 0, 2, seq, 5000, deal); a.renderDevice(p, ptr)
            a.frameEnd(ctrl, 2, seq, 5000)
            sa, sb = a.frameSyncStatus(), b.frameSyncStatus()
            assert (sa, sb) == (0, 0), (seq, sa, sb)
            got = _view_i32(ptr, n_out).cpu().numpy().view(np.uint32).reshape(h >> k, w >> k)
            assert np.array_equal(got, want)
            ca, cb = a.counters()["samples"], b.counters()["samples"]
            print("two GPUs, frame %d: %d + %d samples" % (seq, ca, cb))
            assert ca + cb == w * h
            if seq > 1:
                assert ca > 0 and cb > 0
        a.sharedFree(ptr)


def test_error_message_through_the_context_from_another_thread(sr, R):
    """rmdf_ctx_last_error: the message of a failing call is readable from ANY OS thread (a Haskell `safe` foreign
    call may run the call and the fetch on different threads); the thread-local rmdf_last_error is empty there."""
    import threading
    p = sr.makeParams(R.FragmentShader.FSDETestShader, None, 64, 64, 0.0, resolve_log2=5)
    assert sr._lib.rmdf_render_device(sr.handle, C.byref(p), C.c_void_p(16)) == R._capi.ERR_ARG
    got = {}

    def fetch():
        buf = C.create_string_buffer(512)
        got["rc"] = sr._lib.rmdf_ctx_last_error(sr.handle, buf, 512)
        got["ctx"] = buf.value.decode()
        got["tls"] = sr._lib.rmdf_last_error().decode()
    t = threading.Thread(target=fetch); t.start(); t.join()
    assert got["rc"] == 0 and "resolve_log2" in got["ctx"] and got["tls"] == ""
    # the caller's current device is what it was (DeviceGuard)
    import torch
    assert torch.cuda.current_device() == 0


def test_work_order_never_changes_a_frame(sr, R, oracle_envs):
    """rmdf_set_work_order (frame_sync.cuh "work order"): from the second frame of a sequence on, the kernel claims
    its items most-expensive-first from the previous frame's measured costs.  Pure scheduling — every frame of the
    sequence, with other work lists (another size, a tile, a moved camera) in between, equals the scan-order frame
    word for word, and every sample is rendered exactly once (counters)."""
    import torch
    for s in range(5):
        sr.bindCubeMap(s, oracle_envs[s])
    w, h = 2048, 1536
    out = torch.zeros(w * h, dtype=torch.int32, device="cuda")

    def frame(scene, t, mode, ww=w, hh=h, **kw):
        out.zero_()
        sr.resetCounters()
        sr.renderDevice(sr.makeParams(scene, None, ww, hh, t, mode=mode, **kw), out.data_ptr())
        sr.sync()
        c = sr.counters()
        return out[: (ww >> kw.get("resolve_log2", 0)) * (hh >> kw.get("resolve_log2", 0))].cpu().numpy().copy(), \
            {k: c[k] for k in ("samples", "in_sphere", "hits", "march_de", "shade_de", "inner_iters")}

    L = sr._lib
    L.rmdf_debug_order_get.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.rmdf_debug_order_get_order.argtypes = [C.c_void_p, C.c_void_p, C.c_int]

    def check_order():
        """the permutation the device made = a stable sort of the recorded costs by octave, most expensive first"""
        rec, order = np.zeros(1 << 20, np.uint32), np.zeros(1 << 20, np.uint32)
        n = L.rmdf_debug_order_get(sr.handle, rec.ctypes.data, rec.size)
        assert n > 0 and L.rmdf_debug_order_get_order(sr.handle, order.ctypes.data, order.size) == n
        lg = np.floor(np.log2(np.maximum(rec[:n], 1).astype(np.float64))).astype(np.int64)
        assert np.array_equal(order[:n], np.argsort(-lg, kind="stable").astype(np.uint32))
        return n

    for scene, mode in ((2, R.MODE_PARITY), (2, R.MODE_FAST), (3, R.MODE_FAST)):
        sr.setWorkOrder(False)
        ref0, cnt0 = frame(scene, 0.0, mode)
        ref1, cnt1 = frame(scene, 0.7, mode)
        refk, cntk = frame(scene, 0.0, mode, resolve_log2=2)
        sr.setWorkOrder(True)
        launches = sr.counters()["launches"]
        for i in range(3):                                   # frame 0 records, frames 1 and 2 run in cost order
            got, cnt = frame(scene, 0.0, mode)
            assert np.array_equal(got, ref0) and cnt == cnt0, (scene, mode, i)
            assert check_order() >= 4096
        assert sr.counters()["launches"] == launches + 6     # each frame: the march kernel + the order kernel
        got, cnt = frame(scene, 0.7, mode)                   # the camera moved: last frame's order, still every item once
        assert np.array_equal(got, ref1) and cnt == cnt1
        got, cnt = frame(scene, 0.0, mode, 520, 500)         # another work list in between (ragged edges)
        check_order()
        got, cnt = frame(scene, 0.0, mode, resolve_log2=2)   # and another one: starts over in scan order
        assert np.array_equal(got, refk) and cnt == cntk
        got, cnt = frame(scene, 0.0, mode, resolve_log2=2)
        assert np.array_equal(got, refk) and cnt == cntk
        got, cnt = frame(scene, 0.0, mode)
        assert np.array_equal(got, ref0) and cnt == cnt0
    launches = sr.counters()["launches"]
    frame(0, 0.0, R.MODE_FAST); frame(1, 0.0, R.MODE_FAST)
    assert sr.counters()["launches"] == launches + 2         # Cornell and CSG frames are not ordered (nothing to gain)
