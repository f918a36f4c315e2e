#!/usr/bin/env python
"""A/B timing of kernel variants on one B200 (run under gpurun): for each named setting of the tuning knobs (environment
variables read by srcnn_create) a fresh engine times the fused SRCNN kernel alone (stage_cnn on a 3840x2160 Y plane,
8 rotating plane pairs, CUDA events around every launch, median + mean of 40) and the colour+bicubic kernel alone
(1080p -> 4K), and checks that every setting produces the SAME BYTES as the first one (the knobs must not touch results).
Not a bench value: bench.py measures the whole path."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CONFIGS = [("ovh 12", dict(SRCNN_TC2_SEG_OVH="12"))]
CHECK_BYTES = True   # among the settings that leave the arithmetic on (WHATIF=0)


def main():
    import srcnn_cpp_b200 as S
    OW, OH = 3840, 2160
    rng = np.random.default_rng(0)
    nbuf = 8
    pitch = (OW + 127) // 128 * 128
    ys = [torch.from_numpy(rng.integers(0, 256, (OH, pitch), dtype=np.uint8)).cuda()[:, :OW] for _ in range(nbuf)]
    outs = [torch.zeros((OH, pitch), dtype=torch.uint8, device="cuda")[:, :OW] for _ in range(nbuf)]
    src = torch.from_numpy(rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)).cuda()
    planes = [torch.zeros((OH, pitch), dtype=torch.uint8, device="cuda")[:, :OW] for _ in range(3)]
    ref_bytes = None
    rows = []
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    torch.cuda.set_stream(st)   # the events below are recorded on the stream the kernels are launched on
    for name, env in CONFIGS:
        os.environ.update(env)
        eng = S.Engine(device=0, variant=S.VARIANT_TC)
        eng.set_stream(st.cuda_stream)
        for i in range(5):
            eng.stage_cnn(ys[i % nbuf], outs[i % nbuf], variant=S.VARIANT_TC)
        eng.sync()
        torch.cuda.synchronize()
        ts = []
        for i in range(40):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.stage_cnn(ys[i % nbuf], outs[i % nbuf], variant=S.VARIANT_TC)
            e1.record()
            ts.append((e0, e1))
        eng.sync()
        torch.cuda.synchronize()
        ms = np.array([a.elapsed_time(b) for a, b in ts])
        got = outs[0].cpu().numpy()
        same = True if ref_bytes is None else bool(np.array_equal(ref_bytes, got))
        if ref_bytes is None:
            ref_bytes = got.copy()
        # colour + bicubic alone
        for i in range(3):
            eng.stage_color_bicubic(src, 2.0, *planes)
        eng.sync()
        torch.cuda.synchronize()
        ta = []
        for i in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.stage_color_bicubic(src, 2.0, *planes)
            e1.record()
            ta.append((e0, e1))
        eng.sync()
        torch.cuda.synchronize()
        msa = np.array([a.elapsed_time(b) for a, b in ta])
        eng.close()
        row = dict(config=name, cnn_ms_median=float(np.median(ms)), cnn_ms_mean=float(ms.mean()), cnn_ms_min=float(ms.min()),
                   colour_ms_median=float(np.median(msa)), same_bytes=same)
        rows.append(row)
        print(json.dumps(row), flush=True)
    assert not CHECK_BYTES or all(r["same_bytes"] for r in rows), "a tuning knob changed the results"


if __name__ == "__main__":
    main()
