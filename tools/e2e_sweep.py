#!/usr/bin/env python
"""Sweep of the host-buffer pipeline's sub-band schedule for one 1080p -> 4K frame (srcnn_process_host, pinned buffers, graph replay):
ms per call for (max bands, first-band rows, growth).  Same bytes for every schedule (tests/test_mgpu.py)."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srcnn_cpp_b200 as S
from bench import synth_frame
W, H, OW, OH = 1920, 1080, 3840, 2160
eng = S.Engine(0)
L = eng.L
L.srcnn_debug_set_band_schedule.argtypes = [C.c_void_p, C.c_int, C.c_double]
pin_in, pin_out = S.PinnedBuffer(H * W * 3), S.PinnedBuffer(OH * OW * 3)
pin_in.array[:] = synth_frame(0).reshape(-1)
def run(n):
    t0 = time.perf_counter()
    for _ in range(n):
        assert L.srcnn_process_host(eng.ctx, pin_in.ptr, W, H, 3 * W, S.ORDER_BGR, C.c_float(2.0), pin_out.ptr, 3 * OW) == 0
    return (time.perf_counter() - t0) / n * 1e3
ref = None
for hb, first, growth in [(8, 0, 1.45), (1, 0, 1.45), (2, 0, 1.45), (4, 0, 1.45), (8, 0, 1.0), (8, 128, 1.6), (8, 160, 1.5), (8, 192, 1.5), (8, 192, 1.7), (8, 256, 1.3),
                          (8, 256, 1.6), (8, 256, 2.0), (8, 320, 1.45), (8, 400, 1.3), (12, 128, 1.35), (16, 128, 1.25), (8, 0, 1.45)]:
    eng.set_host_bands(hb)
    L.srcnn_debug_set_band_schedule(eng.ctx, first, growth)
    run(4)
    ms = min(run(30) for _ in range(3))
    out = pin_out.array.copy()
    same = True if ref is None else bool(np.array_equal(ref, out))
    if ref is None:
        ref = out
    print(json.dumps(dict(host_bands=hb, first_rows=first, growth=growth, ms=round(ms, 4), GPix_s=round(OW * OH / ms / 1e6, 2), same_bytes=same)), flush=True)
