"""Host-buffer pipeline (srcnn_process_host / _batch_host / _band_host): CUDA-graph replay of repeated calls.
The second time a call is seen (same pinned buffers, same geometry) its three-stream pipeline is captured into a CUDA graph;
from then on it is one cudaGraphLaunch.  A replay must read the buffers' CURRENT contents and equal the live run."""
import ctypes as C

import numpy as np
import pytest

from conftest import natural_like

pytestmark = pytest.mark.gpu


def _graphs(engine, on=-1):
    engine.L.srcnn_debug_graphs.argtypes = [C.c_void_p, C.c_int]
    return engine.L.srcnn_debug_graphs(engine.ctx, on)


@pytest.mark.parametrize("h,w,scale", [(540, 300, 2.0), (90, 120, 1.5)])
def test_graph_replay_equals_live(engine, h, w, scale):
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(h)
    ow, oh = S.out_dims(w, h, scale)
    pin_in, pin_out = S.PinnedBuffer(h * w * 3), S.PinnedBuffer(oh * ow * 3)
    imgs = [natural_like(rng, h, w) for _ in range(4)]
    _graphs(engine, 0)
    live = [engine.process(im, scale) for im in imgs]          # pageable buffers, graphs off: the live pipeline
    _graphs(engine, 1)
    before = _graphs(engine)
    got = []
    for im in imgs:                                            # same pinned buffers every time: call 1 live, call 2 captured, then replays
        pin_in.array[:] = im.reshape(-1)
        pin_out.array[:] = 0
        rc = engine.L.srcnn_process_host(engine.ctx, pin_in.ptr, w, h, 3 * w, S.ORDER_BGR, C.c_float(scale), pin_out.ptr, 3 * ow)
        assert rc == 0
        got.append(pin_out.array.reshape(oh, ow, 3).copy())
    assert _graphs(engine) == before + 1                       # exactly one pipeline was instantiated
    for a, b in zip(got, live):
        assert np.array_equal(a, b)
    # pageable buffers never take the graph path, and still work after graphs exist
    assert np.array_equal(engine.process(imgs[0], scale), live[0])
    # a different variant is a different key
    engine.set_variant(S.VARIANT_FP32)
    try:
        pin_in.array[:] = imgs[1].reshape(-1)
        for _ in range(3):
            rc = engine.L.srcnn_process_host(engine.ctx, pin_in.ptr, w, h, 3 * w, S.ORDER_BGR, C.c_float(scale), pin_out.ptr, 3 * ow)
            assert rc == 0
        fp = pin_out.array.reshape(oh, ow, 3).copy()
        _graphs(engine, 0)
        assert np.array_equal(fp, engine.process(imgs[1], scale))
    finally:
        engine.set_variant(S.VARIANT_TC)
        _graphs(engine, 1)
    pin_in.free()
    pin_out.free()


def test_launch_count_counts_replays(engine):
    """srcnn_launch_count keeps counting kernels when a call is replayed from a graph (bench.py's gpu_launches)."""
    import srcnn_cpp_b200 as S
    h, w = 64, 96
    ow, oh = S.out_dims(w, h, 2.0)
    pin_in, pin_out = S.PinnedBuffer(h * w * 3), S.PinnedBuffer(oh * ow * 3)
    pin_in.array[:] = 77
    counts = []
    for _ in range(4):
        n0 = engine.launches
        assert engine.L.srcnn_process_host(engine.ctx, pin_in.ptr, w, h, 3 * w, S.ORDER_BGR, C.c_float(2.0), pin_out.ptr, 3 * ow) == 0
        counts.append(engine.launches - n0)
    assert counts[0] == 3 and all(c == counts[0] for c in counts), counts
    pin_in.free()
    pin_out.free()
