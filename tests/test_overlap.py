"""GPU tests of the cross-call overlap: consecutive device-resident whole-path calls alternate between two plane sets and the
colour+bicubic kernel of call i+1 is launched with programmatic stream serialisation, so it may run beside the merge kernel of
call i (srcnn_cpp_b200/csrc/api.cu, may_start_early).  Results must not depend on it -- also when a call's source IS the previous
call's result (the early start is then refused), when calls of different geometry alternate, and for batches.
Reference path replaced: the timed body src/srcnn.cpp:505-659, called frame after frame."""
import ctypes as C

import numpy as np
import pytest

from conftest import natural_like

pytestmark = pytest.mark.gpu


def _overlap(engine, on):
    engine.L.srcnn_debug_overlap.argtypes = [C.c_void_p, C.c_int]
    engine.L.srcnn_debug_overlap.restype = C.c_longlong
    return int(engine.L.srcnn_debug_overlap(engine.ctx, on))


@pytest.mark.parametrize("variant_name", ["tc", "fp32"])
def test_back_to_back_calls_equal_serialised_calls(engine, variant_name):
    import torch
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(5)
    geos = [(320, 200, 2.0), (256, 144, 4.0), (320, 200, 2.0), (200, 120, 1.5), (320, 200, 2.0), (320, 200, 2.0), (256, 144, 4.0), (256, 144, 4.0)]
    imgs = [torch.from_numpy(natural_like(rng, h, w)).to("cuda:0") for (w, h, s) in geos]
    outs = lambda: [torch.zeros((S.out_dims(w, h, s)[1], S.out_dims(w, h, s)[0], 3), dtype=torch.uint8, device="cuda:0") for (w, h, s) in geos]
    engine.set_variant(S.VARIANT_FP32 if variant_name == "fp32" else S.VARIANT_TC)
    try:
        _overlap(engine, 0)
        ref = outs()
        for im, (w, h, s), o in zip(imgs, geos, ref):
            engine.process_device(im, s, o)
            engine.sync()
        early0 = _overlap(engine, 1)
        for rep in range(3):
            got = outs()
            for im, (w, h, s), o in zip(imgs, geos, got):
                engine.process_device(im, s, o)          # no synchronisation in between
            engine.sync()
            for a, b in zip(ref, got):
                assert torch.equal(a, b)
        assert _overlap(engine, -1) - early0 >= 3 * 2     # a call that follows one of the same plane geometry was allowed to start early
    finally:
        _overlap(engine, 1)
        engine.set_variant(S.VARIANT_TC)


def test_result_fed_to_the_next_call_is_not_read_early(engine):
    """x2 of x2: call 2 reads what call 1's merge kernel writes -- it must wait for it."""
    import torch
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(6)
    img = torch.from_numpy(natural_like(rng, 180, 256)).to("cuda:0")
    warm = torch.zeros((360, 512, 3), dtype=torch.uint8, device="cuda:0")
    mid0, out0 = torch.zeros_like(warm), torch.zeros((720, 1024, 3), dtype=torch.uint8, device="cuda:0")
    _overlap(engine, 0)
    engine.process_device(img, 2.0, mid0)
    engine.sync()
    engine.process_device(mid0, 2.0, out0)
    engine.sync()
    _overlap(engine, 1)
    for rep in range(4):
        mid, out = torch.zeros_like(mid0), torch.zeros_like(out0)
        engine.process_device(img, 2.0, warm)             # a previous call whose merge the next one could overlap
        early = _overlap(engine, -1)
        engine.process_device(img, 2.0, mid)
        assert _overlap(engine, -1) == early + 1          # independent of `warm`: early start allowed
        engine.process_device(mid, 2.0, out)
        assert _overlap(engine, -1) == early + 1          # reads `mid`: refused
        engine.sync()
        assert torch.equal(mid, mid0) and torch.equal(out, out0)


def test_batches_back_to_back(engine):
    import torch
    rng = np.random.default_rng(7)
    frames = torch.from_numpy(np.stack([natural_like(rng, 90, 160) for _ in range(6)])).to("cuda:0")
    _overlap(engine, 0)
    ref = torch.zeros((6, 180, 320, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_batch_device(frames, 2.0, ref)
    engine.sync()
    _overlap(engine, 1)
    outs = [torch.zeros_like(ref) for _ in range(4)]
    for o in outs:
        engine.process_batch_device(frames, 2.0, o)
    engine.sync()
    for o in outs:
        assert torch.equal(o, ref)


def test_profile_mode_2_times_the_cnn_stage_and_the_gaps(engine):
    import torch
    rng = np.random.default_rng(8)
    img = torch.from_numpy(natural_like(rng, 270, 480)).to("cuda:0")
    out = [torch.zeros((540, 960, 3), dtype=torch.uint8, device="cuda:0") for _ in range(2)]
    for k in range(3):
        engine.process_device(img, 2.0, out[k & 1])
    engine.profile_enable(2)
    try:
        for k in range(6):
            engine.process_device(img, 2.0, out[k & 1])
        ms, calls = engine.profile_read()
    finally:
        engine.profile_enable(False)
    assert calls == 6 and ms[1] > 0 and ms[0] > 0 and ms[2] == 0
    engine.profile_enable(True)
    try:
        for k in range(6):
            engine.process_device(img, 2.0, out[k & 1])
        ms1, calls1 = engine.profile_read()
    finally:
        engine.profile_enable(False)
    assert calls1 == 6 and min(ms1) > 0


def test_two_contexts_on_one_stream_chained(engine):
    """A second context on the same stream whose source is the first context's result: the early start is refused (the second
    context cannot know what the first one's merge writes)."""
    import torch
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(9)
    img = torch.from_numpy(natural_like(rng, 180, 256)).to("cuda:0")
    eng2 = S.Engine(device=0, variant=S.VARIANT_TC, stream=torch.cuda.current_stream().cuda_stream)   # the fixture's stream
    try:
        mid0 = torch.zeros((360, 512, 3), dtype=torch.uint8, device="cuda:0")
        out0 = torch.zeros((720, 1024, 3), dtype=torch.uint8, device="cuda:0")
        _overlap(engine, 0)
        engine.process_device(img, 2.0, mid0)
        engine.sync()
        eng2.process_device(mid0, 2.0, out0)
        eng2.sync()
        _overlap(engine, 1)
        other = torch.zeros((720, 1024, 3), dtype=torch.uint8, device="cuda:0")
        for rep in range(4):
            mid, out = torch.zeros_like(mid0), torch.zeros_like(out0)
            eng2.process_device(mid0, 2.0, other)     # eng2's own previous call: same geometry, unrelated buffers
            engine.process_device(img, 2.0, mid)
            eng2.process_device(mid, 2.0, out)        # reads what engine's merge is writing
            eng2.sync()
            assert torch.equal(mid, mid0) and torch.equal(out, out0)
    finally:
        eng2.close()
