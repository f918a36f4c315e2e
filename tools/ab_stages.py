#!/usr/bin/env python
"""Same-box timing of the two HBM-bound stages alone (run under gpurun): colour+bicubic and merge+colour-back, each over 8
rotating source / plane / result sets (249 MB > the 126 MB L2, like bench.py's step), CUDA events on the launching stream,
median of 40.  usage: ab_stages.py [W H SCALE]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srcnn_cpp_b200 as S

W, H, SC = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 2.0)
OW, OH = S.out_dims(W, H, SC)
NB = 8
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
eng = S.Engine(0, stream=st.cuda_stream)
rng = np.random.default_rng(0)
srcs = [torch.from_numpy(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).cuda() for _ in range(NB)]
pitch = (OW + 127) // 128 * 128
planes = [[torch.zeros((OH, pitch), dtype=torch.uint8, device="cuda")[:, :OW] for _ in range(3)] for _ in range(NB)]
dsts = [torch.zeros((OH, OW, 3), dtype=torch.uint8, device="cuda") for _ in range(NB)]


def timed(fn, n=40, warm=5):
    for i in range(warm):
        fn(i)
    eng.sync()
    ev = []
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        ev.append((e0, e1))
    eng.sync()
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    return float(np.median(ms)), float(ms.min())


px = OW * OH
med, mn = timed(lambda i: eng.stage_color_bicubic(srcs[i % NB], SC, *planes[i % NB]))
gbs = (3.0 / (SC * SC) + 3.0) * px / (med * 1e-3) / 1e9
print(json.dumps(dict(stage="colour_bicubic", ms_median=med, ms_min=mn, GBs=gbs)), flush=True)
med, mn = timed(lambda i: eng.stage_merge(planes[i % NB][0], planes[i % NB][1], planes[i % NB][2], dsts[i % NB]))
print(json.dumps(dict(stage="merge", ms_median=med, ms_min=mn, GBs=6.0 * px / (med * 1e-3) / 1e9)), flush=True)
eng.close()
