#!/usr/bin/env python
"""Throughput of the JPEG stream ingest path (srcnn_jpeg_stream_process) on one GPU: frames/s and output MPix/s, wall clock
around the call (decode + kernels + encode, bitstreams in host memory on both sides).  usage: stream_bench.py W H SCALE NFRAMES"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import srcnn_cpp_b200 as S
from bench import synth_frame

W, H, SC, N = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080, 2.0, 64)
eng = S.Engine(0)
st = S.JpegStream(eng, 95)
base = [cv2.imencode(".jpg", synth_frame(k, H, W), [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes() for k in range(4)]
jpegs = [base[k % 4] for k in range(N)]
st.process(jpegs[:4], SC)
t0 = time.perf_counter()
outs, (ow, oh) = st.process(jpegs, SC)
dt = time.perf_counter() - t0
print(json.dumps(dict(what="JPEG stream %dx%d -> %dx%d x%g, %d frames, decode + kernels + encode" % (W, H, ow, oh, SC, N), frames_per_s=N / dt,
                      out_MPix_per_s=N * ow * oh / dt / 1e6, ms_per_frame=dt / N * 1e3, in_MB=sum(map(len, jpegs)) / 1e6,
                      out_MB=sum(map(len, outs)) / 1e6)), flush=True)
st.close()
eng.close()
