"""The multi-GPU driver behind the C ABI (srcnn_mgpu_*, include/srcnn_b200.h; SURVEY 8b "device list", 8e).

CPU: the band plan (pure host arithmetic) and the loud failure without a device.
GPU: frames and row bands fanned out over several workers in ONE process equal the one-context result bit for
bit, with the oracle checked at every band seam.  A worker list may name a device more than once, so the threading,
the seams and the strided frame split are exercised on a one-GPU box too; with >= 2 GPUs the same tests run on
distinct devices."""
import ctypes as C

import numpy as np
import pytest

from conftest import diff_stats, natural_like


def test_band_plan_covers_every_row_once_and_brings_the_halo():
    import srcnn_cpp_b200 as S
    for h, scale, n in [(1080, 2.0, 8), (720, 2.0, 3), (2160, 4.0, 8), (384, 1.5, 5), (32768, 2.0, 8), (5, 2.0, 8)]:
        ow, oh = S.out_dims(16, h, scale)
        nxt = 0
        for i in range(n):
            r0, r1, s0, s1 = S.mgpu_band_plan(n, h, scale, i)
            assert r0 == nxt and r1 >= r0
            nxt = r1
            if r1 > r0:
                assert (s0, s1) == S.band_src_rows(h, scale, r0, r1)
                assert 0 <= s0 < s1 <= h
        assert nxt == oh
    L = S.load_library()
    v = [C.c_int() for _ in range(4)]
    assert L.srcnn_mgpu_band_plan(0, 100, C.c_float(2.0), 0, *[C.byref(x) for x in v]) == S.E_ARG
    assert L.srcnn_mgpu_band_plan(4, 100, C.c_float(2.0), 4, *[C.byref(x) for x in v]) == S.E_ARG
    assert L.srcnn_mgpu_band_plan(4, 100, C.c_float(-1.0), 0, *[C.byref(x) for x in v]) == S.E_RATIO


def test_mgpu_without_a_device_fails_loudly():
    import torch
    import srcnn_cpp_b200 as S
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(S.SrcnnError) as e:
        S.MultiEngine([0, 1])
    assert e.value.status == S.E_NODEVICE


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0, 0]]                       # three workers on one GPU: threads, seams and frame strides without a second GPU
    if n >= 2:
        lists.append(list(range(min(n, 8))))  # one worker per real device
    return lists


@pytest.mark.gpu
@pytest.mark.parametrize("variant_name", ["fp32", "tc"])
def test_banded_host_equals_one_gpu_and_oracle_at_seams(engine, oracle, variant_name):
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(11)
    h, w, scale = 150, 200, 2.0
    img = natural_like(rng, h, w)
    variant = S.VARIANT_FP32 if variant_name == "fp32" else S.VARIANT_TC
    engine.set_variant(variant)
    try:
        whole = engine.process(img, scale)
    finally:
        engine.set_variant(S.VARIANT_TC)
    want = oracle.pipeline(img, scale)
    for devs in _device_lists():
        m = S.MultiEngine(devs, variant)
        try:
            got = m.process_banded(img, scale)
            ms, wall = m.last_timing()
            plan = m.band_plan(h, scale)
        finally:
            m.close()
        assert np.array_equal(got, whole), devs               # bands over workers == one context, bit for bit
        assert len(ms) == len(devs) and wall > 0 and all(t > 0 for t in ms)
        for (r0, r1, _, _) in plan[1:]:                          # 16 rows around every seam against the oracle
            a, b = max(r0 - 8, 0), min(r0 + 8, got.shape[0])
            st = diff_stats(got[a:b], want[a:b])
            if variant_name == "fp32":
                assert st["max"] == 0, (devs, r0, st)
            else:
                assert st["max"] <= 2 and st["le1"] >= 0.999, (devs, r0, st)


@pytest.mark.gpu
def test_batch_host_frames_round_robin(engine):
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(12)
    frames = np.stack([natural_like(rng, 72, 96) for _ in range(7)])      # 7 frames over 3 workers: shares of 3, 2, 2
    want = engine.process_batch(frames, 2.0)
    for devs in _device_lists():
        m = S.MultiEngine(devs)
        try:
            got = m.process_batch(frames, 2.0)
            assert np.array_equal(got, want), devs
            one = m.process_batch(frames[:1], 2.0)                         # fewer frames than workers
            assert np.array_equal(one, want[:1])
        finally:
            m.close()


@pytest.mark.gpu
def test_device_resident_shares(engine):
    """srcnn_mgpu_process_batch_device / _banded_device: every worker's share lives on its own device."""
    import torch
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(13)
    h, w, scale = 120, 136, 2.0
    ow, oh = S.out_dims(w, h, scale)
    frames = np.stack([natural_like(rng, h, w) for _ in range(5)])
    want = engine.process_batch(frames, scale)
    for devs in _device_lists():
        m = S.MultiEngine(devs)
        try:
            n = m.n
            srcs, dsts = [], []
            for i, d in enumerate(m.devices):
                mine = frames[i::n]
                srcs.append(torch.from_numpy(np.ascontiguousarray(mine)).to("cuda:%d" % d) if len(mine) else None)
                dsts.append(torch.zeros((len(mine), oh, ow, 3), dtype=torch.uint8, device="cuda:%d" % d) if len(mine) else None)
            torch.cuda.synchronize()
            m.process_batch_device(srcs, scale, dsts)
            for i in range(n):
                if dsts[i] is not None:
                    assert np.array_equal(dsts[i].cpu().numpy(), want[i::n]), (devs, i)
            # one image, one band per worker
            plan = m.band_plan(h, scale)
            sb = [torch.from_numpy(np.ascontiguousarray(frames[0][s0:s1])).to("cuda:%d" % d) for (r0, r1, s0, s1), d in zip(plan, m.devices)]
            db = [torch.zeros((r1 - r0, ow, 3), dtype=torch.uint8, device="cuda:%d" % d) for (r0, r1, s0, s1), d in zip(plan, m.devices)]
            torch.cuda.synchronize()
            m.process_banded_device(sb, w, h, scale, db)
            got = np.concatenate([t.cpu().numpy() for t in db])
            assert np.array_equal(got, want[0]), devs
            ms, wall = m.last_timing()
            assert all(t > 0 for t in ms)
        finally:
            m.close()


@pytest.mark.gpu
def test_calls_leave_the_callers_device_alone(engine):
    """Every entry point restores the calling thread's current CUDA device (a host application keeps its own)."""
    import torch
    import srcnn_cpp_b200 as S
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    torch.cuda.set_device(1)
    try:
        eng = S.Engine(device=0)
        img = np.zeros((20, 20, 3), np.uint8)
        eng.process(img, 2.0)
        eng.close()
        assert torch.cuda.current_device() == 1
        x = torch.zeros(4, device="cuda")
        assert x.device.index == 1
    finally:
        torch.cuda.set_device(0)


@pytest.mark.gpu
@pytest.mark.parametrize("scale,r0,r1", [(2.0, 0, 37), (2.0, 100, 260), (3.0, 200, 450), (1.5, 7, 8)])
def test_band_host_equals_rows_of_the_whole(engine, scale, r0, r1):
    """srcnn_process_band_host: rows [r0, r1) computed from host buffers (only the needed source rows are copied in)."""
    rng = np.random.default_rng(int(scale * 10) + r0)
    img = natural_like(rng, 150, 90)
    whole = engine.process(img, scale)
    r1 = min(r1, whole.shape[0])
    got = engine.process_band_host(img, scale, r0, r1)
    assert np.array_equal(got, whole[r0:r1])


@pytest.mark.gpu
def test_tall_single_frame_is_cut_into_sub_bands(engine):
    """The host pipeline cuts a tall frame into sub-bands (>= 1024 output rows); any sub-band count gives the same bytes."""
    rng = np.random.default_rng(5)
    img = natural_like(rng, 700, 64)
    outs = []
    for bands in (1, 2, 5):
        engine.set_host_bands(bands)
        outs.append(engine.process(img, 2.0))
    engine.set_host_bands(8)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [2.0, 1.5, 4.0, 0.5])
def test_resize_plane_is_cv_cubic(engine, oracle, scale):
    """srcnn_resize_plane_host == resize(plane, CV_INTER_CUBIC) of the reference path (src/srcnn.cpp:577-582)."""
    rng = np.random.default_rng(int(scale * 8))
    plane = rng.integers(0, 256, (45, 70), dtype=np.uint8)
    got = engine.resize_plane(plane, scale)
    want = oracle.resize_cubic(plane, got.shape[1], got.shape[0])
    assert np.array_equal(got, want)
