#!/usr/bin/env python
"""bench.py -- headline benchmark of the SRCNN hot path (BASELINE.json: output MPix/s, 1080p -> 4K x2).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA library through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation
  python bench.py --config cfg3|cfg4|cfg5 --gpus N ...     the other BASELINE configurations (strong scaling)
  python bench.py --config cfg3 --gpus N --mgpu            the same through the in-library multi-GPU driver (one process)
  python bench.py --sustain S                              timed region of at least S seconds (power steady state)

One "step" = one pass of the whole hot path (colour+bicubic -> fused SRCNN -> merge+colour back) over one batch of
synthetic input:
  cfg2 (default, BASELINE configs[1]): one 1920x1080 BGR frame per rank -> 3840x2160.  With N > 1 every rank processes its
        own frames (frame sharding, SURVEY 8e: no collective on the data path) -> WEAK scaling.
  cfg3 (configs[2]): 1024 frames of 1280x720 -> 2560x1440, frame f on rank f mod N            -> STRONG scaling
  cfg4 (configs[3]): one 32768x32768 -> 65536x65536 image in N row bands with the 6-px halo   -> STRONG scaling
  cfg5 (configs[4]): a stream of 64 frames 3840x2160 -> 15360x8640 (x4), frame f on rank f mod N -> STRONG scaling
`value` = all ranks' output pixels / max-over-ranks device time.

  value      : inputs resident in HBM, CUDA-event timed on the launching stream, K steps back to back; cfg2 rotates over
               more frame/result buffers than fit in L2, the other configurations are far larger than L2 per step
  e2e        : the same metric through the host-buffer C-ABI calls with pinned HOST buffers: H2D of the input and D2H of
               the result are inside the timed region of every step
  roofline   : fused SRCNN kernel, 16 064 algorithmic FLOP per output pixel / its CUDA-event duration (events around every
               launch of the timed region, on the launching stream), against MEASURED_PEAKS.json's dense bf16 burst figure
               (same tensor rate as fp16); the sustained figure beside it
  stages     : the two HBM-bound kernels are timed in a short SERIALISED pass after the timed region (events around every
               stage): in the timed region the merge of step i and the colour+bicubic of step i+1 run side by side (the
               library's cross-call overlap), and an event between them would separate them; `between_srcnn_launches_ms` is
               what the pair costs there
  cpu_baseline: the reference's CPU code (oracle/_ref/libref.so + cv2, IPP off) on the box's host cores, on a bounded
               sample of the same workload
"""
import argparse
import csv
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PX = 16064           # conv1 10368 + conv2 4096 + conv3 1600
UNIT = "MPix/s"

# name -> source geometry, scale, frames per step (whole job), scaling, BASELINE wording
CONFIGS = {
    "cfg2": dict(w=1920, h=1080, scale=2.0, frames=1, scaling="weak",
                 metric="output MPix/s (1080p->4K x2)",
                 workload="single 1920x1080 -> 3840x2160 x2 synthetic BGR frame per rank per step (BASELINE configs[1])"),
    "cfg3": dict(w=1280, h=720, scale=2.0, frames=1024, scaling="strong",
                 metric="output MPix/s (1024 x 720p->1440p x2)",
                 workload="batch of 1024 synthetic 1280x720 frames x2 per step, frame f on GPU f mod N (BASELINE configs[2])"),
    "cfg4": dict(w=32768, h=32768, scale=2.0, frames=1, scaling="strong",
                 metric="output MPix/s (32768^2->65536^2 x2, row bands)",
                 workload="one synthetic 32768x32768 image x2 per step, N row bands with the 6-px halo, band i on GPU i (BASELINE configs[3])"),
    "cfg5": dict(w=3840, h=2160, scale=4.0, frames=64, scaling="strong",
                 metric="output MPix/s (4K->16K x4 frame stream)",
                 workload="stream of 64 synthetic 3840x2160 -> 15360x8640 x4 frames per step, frame f on GPU f mod N (BASELINE configs[4])"),
}


def synth_frame(seed, h=1080, w=1920):
    """Seeded synthetic BGR frame: smooth structure + grain (natural-like statistics, full 0..255 range)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.empty((h, w, 3), np.float32)
    for c in range(3):
        fx, fy, ph = rng.uniform(0.002, 0.02, 3)
        img[:, :, c] = 127 + 90 * np.sin(xx * fx * 6.28 + ph * 100) * np.cos(yy * fy * 6.28) + 30 * np.sin((xx + yy) * 0.05 * (c + 1))
    img += rng.normal(0, 12, img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_device(torch, dev, seed, shape):
    """The same kind of content generated on the device (for inputs too large to build on the host): low-resolution noise
    repeated 64x in both directions plus fine grain.  shape = (..., H, W, 3)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    *lead, h, w, _ = shape
    base = torch.randint(0, 232, (*lead, (h + 63) // 64, (w + 63) // 64, 3), dtype=torch.uint8, device=dev, generator=g)
    out = base.repeat_interleave(64, -3).repeat_interleave(64, -2)[..., :h, :w, :].contiguous()
    del base
    out += torch.randint(0, 24, out.shape, dtype=torch.uint8, device=dev, generator=g)
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=float(d["bf16_tflops"]), bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), src="measured (MEASURED_PEAKS.json, burst)")
    return dict(bf16=1590.0, bf16_sustained=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the fused kernel, read from the newest committed
    `ncu --set full` capture under profiles/ (raw-page CSV: header row, unit row, value row).  -> (bytes, file) or (None, None)."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    pdir = os.path.join(ROOT, "profiles")
    for name in ("r2_k_srcnn_tc_raw.csv", "r2_k_srcnn_tc2_raw.csv", "r1_k_srcnn_tc2_raw.csv"):
        path = os.path.join(pdir, name)
        if not os.path.exists(path):
            continue
        try:
            rows = list(csv.reader(open(path)))
            hdr, units, vals = rows[0], rows[1], rows[2]
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(key)
                tot += float(vals[i].replace(",", "")) * scale[units[i]]
            return tot, "profiles/" + name
        except Exception:
            continue
    return None, None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz, self.power = index, False, [], set(), None, []

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": (max(self.power) if self.power else None)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref when present, else the oracle port) -- the only place bench.py
# executes anything under oracle/.
# ------------------------------------------------------------------------------------------------
def cpu_reference_runner(scale):
    from oracle.oracle import Oracle, RefLib, cv2_pipeline
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
    except Exception:
        pass
    try:
        ref = RefLib("O3")
        kind, threads = "reference", ref.threads()

        def run(img):
            return cv2_pipeline(img, scale, ref.cnn)
        desc = "reference src/srcnn.cpp conv functions compiled unmodified (-O3, bit-identical to the Makefile's -O0) + cv2 (IPP off) stages"
    except Exception:
        orc = Oracle()
        kind, threads = "port", orc.threads()

        def run(img):
            return orc.pipeline(img, scale)
        desc = "oracle/srcnn_oracle.c port (OpenMP)"
    return run, kind, threads, desc


def time_cpu(run, cfg, budget_s=12.0):
    """Times the CPU path on a bounded sample of the workload: a centred crop of one source frame (at most 1920x1080 of it)
    sized so one pass takes roughly budget_s.  Returns (MPix/s, sample description, seconds)."""
    H, W = min(cfg["h"], 1080), min(cfg["w"], 1920)
    s = cfg["scale"]
    frame = synth_frame(0, H, W)
    probe = np.ascontiguousarray(frame[:135, :240])
    run(probe)
    t = time.perf_counter(); out = run(probe); dt = time.perf_counter() - t
    rate = (out.shape[0] * out.shape[1]) / dt                       # output px/s on the probe
    full_px = int(W * s) * int(H * s)
    want_px = min(full_px, max(out.shape[0] * out.shape[1], rate * budget_s))
    frac = (want_px / full_px) ** 0.5
    h = max(135, min(H, int(H * frac) // 2 * 2)); w = max(240, min(W, int(W * frac) // 2 * 2))
    y0, x0 = (H - h) // 2, (W - w) // 2
    crop = np.ascontiguousarray(frame[y0:y0 + h, x0:x0 + w])
    t = time.perf_counter(); out = run(crop); dt = time.perf_counter() - t
    px = out.shape[0] * out.shape[1]
    return px / dt / 1e6, "%dx%d crop of one %dx%d source frame -> %dx%d (%.2f MPix out, %.1f s)" % (w, h, cfg["w"], cfg["h"], out.shape[1], out.shape[0], px / 1e6, dt), dt


def time_cpu_makefile_flags(cfg, budget_s=4.0):
    """The same CPU code as the reference's own Makefile really builds it (no -O flag on the compile lines, Makefile:21-23,43):
    oracle/_ref/libref_O0.so, bit-identical output, ~10x slower.  A small bounded sample; None when that build is absent."""
    try:
        from oracle.oracle import RefLib, cv2_pipeline
        ref = RefLib("O0")
    except Exception:
        return None

    def run(img):
        return cv2_pipeline(img, cfg["scale"], ref.cnn)
    v, sample, _ = time_cpu(run, cfg, budget_s)
    return {"value": v, "unit": UNIT, "cores": ref.threads(), "sample": sample,
            "what": "same sources with the reference Makefile's own compile flags (-mtune=native -fopenmp, no -O)"}


def run_reference_arm(args, cfg, rank, world):
    if rank != 0:
        return
    run, kind, threads, desc = cpu_reference_runner(cfg["scale"])
    run(np.ascontiguousarray(synth_frame(0, 64, 64)))
    per_step_budget = max(1.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    vals, sample = [], ""
    for i in range(args.warmup + args.steps):
        v, sample, dt = time_cpu(run, cfg, per_step_budget)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals]) * 1e3)
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "note": "CPU reference arm: each step is a bounded sample of the workload; rank 0 only"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "what": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm: per-configuration work of one rank.  Each builder returns
#   step(i)            -> enqueues one device-resident step of this rank on the engine's stream
#   e2e_step(i)        -> one host-buffer step (blocking), or None
#   px_step            -> output pixels of the WHOLE job per step
#   e2e_px, h2d, d2h   -> output pixels / bytes one e2e step of this rank moves
#   notes              -> strings for config
# ------------------------------------------------------------------------------------------------
def build_cfg2(S, torch, eng, dev, rank, world, cfg):
    W, H, s = cfg["w"], cfg["h"], cfg["scale"]
    OW, OH = S.out_dims(W, H, s)
    NBUF = 8   # more frame/result buffers than L2 can hold: NBUF * (6.2 MB in + 24.9 MB out) = 249 MB > 126 MB
    frames_h = [synth_frame(1000 * rank + i, H, W) for i in range(2)]
    src = [torch.from_numpy(frames_h[i % 2]).to(dev) for i in range(NBUF)]
    dst = [torch.empty((OH, OW, 3), dtype=torch.uint8, device=dev) for _ in range(NBUF)]
    pin_in = [S.PinnedBuffer(H * W * 3) for _ in range(2)]
    pin_out = [S.PinnedBuffer(OH * OW * 3) for _ in range(2)]
    for k in range(2):
        pin_in[k].array[:] = frames_h[k].reshape(-1)
    L = eng.L

    def step(i):
        eng.process_device(src[i % NBUF], s, dst[i % NBUF])

    def e2e_step(i):
        rc = L.srcnn_process_host(eng.ctx, pin_in[i % 2].ptr, W, H, 3 * W, S.ORDER_BGR, s, pin_out[i % 2].ptr, 3 * OW)
        if rc != 0:
            raise S.SrcnnError(rc, L.srcnn_last_error(eng.ctx).decode())

    def check():
        return int(dst[0][::97, ::101].to(torch.int32).sum().item())
    return dict(step=step, e2e_step=e2e_step, px_step=world * OW * OH, px_rank=OW * OH, e2e_px=world * OW * OH, h2d=H * W * 3, d2h=OH * OW * 3,
                check=check, frames_rank=1, keep=(src, dst, pin_in, pin_out),
                l2="rotating %d frame/result buffers (%.0f MB) > 126 MB L2; no flush kernel in the timed region" % (NBUF, NBUF * (H * W * 3 + OH * OW * 3) / 1e6),
                e2e_api="srcnn_process_host (pinned host buffers, blocking)", e2e_note=None)


def build_frames(S, torch, eng, dev, rank, world, cfg, out_ring=None, e2e_frames=None):
    """cfg3 / cfg5: `frames` frames per step over the whole job, frame f on rank f mod N.  The rank's frames are resident; results
    go to one buffer per frame (cfg3) or rotate over a ring of `out_ring` buffers (cfg5: 398 MB per frame)."""
    W, H, s, N = cfg["w"], cfg["h"], cfg["scale"], cfg["frames"]
    OW, OH = S.out_dims(W, H, s)
    mine = len(range(rank, N, world))
    src = synth_device(torch, dev, 77 + rank, (mine, H, W, 3))
    nout = mine if out_ring is None else min(out_ring, mine)
    dst = torch.empty((nout, OH, OW, 3), dtype=torch.uint8, device=dev)
    L = eng.L

    def step(i):
        if nout == mine:
            eng.process_batch_device(src, s, dst)
        else:
            for f0 in range(0, mine, nout):
                m = min(nout, mine - f0)
                eng.process_batch_device(src[f0:f0 + m], s, dst[:m])

    # e2e on a bounded number of this rank's frames per step (pinned memory for a whole cfg3 share would be 14 GB)
    ef = min(mine, e2e_frames or mine)
    pin_in = S.PinnedBuffer(ef * H * W * 3)
    pin_out = S.PinnedBuffer(ef * OH * OW * 3)
    pin_in.array[:] = src[:ef].reshape(-1).cpu().numpy()

    def e2e_step(i):
        rc = L.srcnn_process_batch_host(eng.ctx, pin_in.ptr, ef, W, H, 3 * W, 3 * W * H, S.ORDER_BGR, s, pin_out.ptr, 3 * OW, 3 * OW * OH)
        if rc != 0:
            raise S.SrcnnError(rc, L.srcnn_last_error(eng.ctx).decode())

    def check():
        return int(dst[0][::97, ::101].to(torch.int32).sum().item())
    return dict(step=step, e2e_step=e2e_step, px_step=N * OW * OH, px_rank=mine * OW * OH, e2e_px=world * ef * OW * OH, h2d=ef * H * W * 3,
                d2h=ef * OH * OW * 3, check=check, frames_rank=mine, keep=(src, dst, pin_in, pin_out),
                l2="%.1f GB of frames + results per rank and step, far beyond the 126 MB L2" % ((mine * H * W * 3 + nout * OH * OW * 3) / 1e9),
                e2e_api="srcnn_process_batch_host (pinned host buffers, H2D / kernels / D2H pipelined over three streams)",
                e2e_note="e2e step = %d of the rank's %d frames" % (ef, mine))


def build_cfg4(S, torch, eng, dev, rank, world, cfg, e2e=True):
    """One gigapixel image in `world` row bands: this rank holds only the source rows its band needs (6-px halo)."""
    W, H, s = cfg["w"], cfg["h"], cfg["scale"]
    OW, OH = S.out_dims(W, H, s)
    r0, r1, s0, s1 = S.mgpu_band_plan(world, H, s, rank)
    src = synth_device(torch, dev, 99, (s1 - s0, W, 3))      # content differs per rank; the timing does not care
    dst = torch.empty((r1 - r0, OW, 3), dtype=torch.uint8, device=dev)
    L = eng.L

    def step(i):
        eng.process_band_device(src, W, H, s0, s1, s, r0, r1, dst)

    e2e_step, pins, h2d, d2h, e2e_px, note = None, None, 0, 0, 0, "e2e skipped"
    if e2e:
        # a bounded band of this rank's share through the host call: 4096 output rows (805 MB out, 201 MB in)
        er1 = min(r1, r0 + 4096)
        es0, es1 = S.band_src_rows(H, s, r0, er1)
        pin_in = S.PinnedBuffer((es1 - es0) * W * 3)
        pin_out = S.PinnedBuffer((er1 - r0) * OW * 3)
        pin_in.array[:] = src[es0 - s0:es1 - s0].reshape(-1).cpu().numpy()
        base = pin_in.ptr - es0 * W * 3        # the call wants the address of source row 0; it only reads rows [es0, es1)
        pins = (pin_in, pin_out)

        def e2e_step(i):
            rc = L.srcnn_process_band_host(eng.ctx, base, W, H, 3 * W, S.ORDER_BGR, s, r0, er1, pin_out.ptr, 3 * OW)
            if rc != 0:
                raise S.SrcnnError(rc, L.srcnn_last_error(eng.ctx).decode())
        h2d, d2h, e2e_px = (es1 - es0) * W * 3, (er1 - r0) * OW * 3, world * (er1 - r0) * OW
        note = "e2e step = output rows [%d, %d) of the rank's band [%d, %d)" % (r0, er1, r0, r1)

    def check():
        return int(dst[::997, ::1013].to(torch.int32).sum().item())
    return dict(step=step, e2e_step=e2e_step, px_step=OW * OH, px_rank=(r1 - r0) * OW, e2e_px=e2e_px, h2d=h2d, d2h=d2h, check=check,
                frames_rank=1, keep=(src, dst, pins),
                l2="%.1f GB of source rows + %.1f GB of result per rank and step, far beyond the 126 MB L2" % (src.numel() / 1e9, dst.numel() / 1e9),
                e2e_api="srcnn_process_band_host (pinned host buffers, sub-bands pipelined over three streams)", e2e_note=note)


def run_ours(args, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import srcnn_cpp_b200 as S

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # an explicit non-default stream: the library launches on it and torch's events are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = S.Engine(device=local_rank, variant=S.VARIANT_FP32 if args.variant == "fp32" else S.VARIANT_TC,
                   stream=stream.cuda_stream)
    if args.config == "cfg2":
        wk = build_cfg2(S, torch, eng, dev, rank, world, cfg)
    elif args.config == "cfg3":
        wk = build_frames(S, torch, eng, dev, rank, world, cfg, e2e_frames=32)
    elif args.config == "cfg5":
        wk = build_frames(S, torch, eng, dev, rank, world, cfg, out_ring=4, e2e_frames=4)
    else:
        wk = build_cfg4(S, torch, eng, dev, rank, world, cfg)
    step, e2e_step = wk["step"], wk["e2e_step"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    steps = args.steps
    if args.sustain > 0:   # as many steps as fill `sustain` seconds (pilot: the warm-up's rate)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for i in range(3):
            step(i)
        p1.record(stream)
        torch.cuda.synchronize()
        steps = max(steps, int(args.sustain * 1e3 / max(1e-3, p0.elapsed_time(p1) / 3)) + 1)
        if world > 1:
            t = torch.tensor([steps], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            steps = int(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    # events around the fused SRCNN kernel only: an event between the merge of step i and the colour+bicubic of step i+1 would
    # separate two kernels the library lets run side by side (cross-call overlap); the other two stages are timed on their own
    # in a short serialised pass after the timed region
    eng.profile_enable(2)
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall = time.perf_counter()
    e0.record(stream)
    for i in range(steps):
        step(args.warmup + i)
    e1.record(stream)
    barrier()
    timed_region_s = time.perf_counter() - t_wall
    ms_total = e0.elapsed_time(e1)
    launches = eng.launches - launches0
    stage_ms, calls = eng.profile_read()
    between_ms = stage_ms[0]          # summed time between consecutive fused-kernel launches (steps - 1 intervals when a step is one launch)
    # serialised pass: every stage bracketed by its own events (no overlap between calls) -> colour+bicubic and merge on their own
    ser_steps = max(1, min(steps, 10 if args.config == "cfg2" else 1))
    eng.profile_enable(True)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for i in range(ser_steps):
        step(args.warmup + i)
    s1.record(stream)
    ser_stage_ms, _ = eng.profile_read()
    ser_ms_step = s0.elapsed_time(s1) / ser_steps
    eng.profile_enable(False)
    stage_ms = [ser_stage_ms[0] * steps / ser_steps, stage_ms[1], ser_stage_ms[2] * steps / ser_steps]

    # ---- end to end through the host-buffer C-ABI call: pinned host memory, H2D + D2H inside ----
    e2e_ms_step = float("nan")
    if e2e_step is not None:
        e2e_steps = max(3, min(steps, 20)) if args.config == "cfg2" else 3
        for i in range(3 if args.config == "cfg2" else 1):
            e2e_step(i)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
        f1.record(stream)
        torch.cuda.synchronize()
        e2e_ms_wall = (time.perf_counter() - t0) * 1e3
        e2e_ms_step = max(f0.elapsed_time(f1), e2e_ms_wall) / e2e_steps   # the call blocks until dst is ready: wall >= device
        barrier()
    # ---- the copies alone: the same bytes per step over PCIe, H2D and D2H concurrently on two streams, all ranks at once.
    # No kernel can make the host-buffer call faster than this (the e2e ceiling of THIS box with THIS many GPUs busy).
    copy_ms_step = float("nan")
    if e2e_step is not None and wk["h2d"] > 0:
        hin, hout = torch.empty(wk["h2d"], dtype=torch.uint8).pin_memory(), torch.empty(wk["d2h"], dtype=torch.uint8).pin_memory()
        din, dout = torch.empty(wk["h2d"], dtype=torch.uint8, device=dev), torch.empty(wk["d2h"], dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        reps = 10 if wk["d2h"] < (1 << 28) else 3

        def copies(k):
            for _ in range(k):
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
        copies(2)
        barrier()
        t0 = time.perf_counter()
        copies(reps)
        torch.cuda.synchronize()
        copy_ms_step = (time.perf_counter() - t0) * 1e3 / reps
        barrier()
        del hin, hout, din, dout
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # sanity: the timed path produced a plausible result (not all zeros) -- cheap guard against silent no-ops
    assert wk["check"]() > 0

    # max over ranks
    t = torch.tensor([ms_total, e2e_ms_step, stage_ms[0], stage_ms[1], stage_ms[2], copy_ms_step, between_ms, ser_ms_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_step, a_ms, b_ms, c_ms, copy_ms_step, between_ms, ser_ms_step = [float(x) for x in t.tolist()]

    if rank == 0:
        peaks = load_peaks()
        s = cfg["scale"]
        ms_step = ms_total / steps
        value = wk["px_step"] / (ms_step * 1e-3) / 1e6
        e2e_value = wk["e2e_px"] / (e2e_ms_step * 1e-3) / 1e6 if e2e_step is not None else None
        # stage times are summed over the timed region: per step of this rank (rank 0's share stands for the others')
        px_rank = wk["px_rank"]
        k_ms = b_ms / steps                                   # fused SRCNN kernel, all launches of one step of a rank
        achieved = FLOP_PER_PX * px_rank / (k_ms * 1e-3) / 1e12
        a_bytes = (3.0 / (s * s) + 3.0) * px_rank             # colour+bicubic: 3/s^2 read + 3 written per output px
        a_gbs = a_bytes / ((a_ms / steps) * 1e-3) / 1e9
        fused_merge = c_ms / steps < 5e-3                     # merge + colour-back runs inside the fused kernel: no K-C launch
        c_gbs = 0.0 if fused_merge else 6.0 * px_rank / ((c_ms / steps) * 1e-3) / 1e9
        traffic, traffic_file = ncu_traffic() if (args.variant != "fp32" and args.config == "cfg2") else (None, None)
        launches_kb = max(1, round(launches / steps / 3)) if args.variant != "fp32" else None
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": "f32 (strict)" if args.variant == "fp32" else "f16 operands, f32 accumulate (tcgen05)",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "variant": args.variant,
                       "parallelism": ("frame-sharded x%d" if args.config != "cfg4" else "row bands x%d") % world + ", no collective",
                       "l2": wk["l2"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": wk["h2d"], "d2h_bytes_per_step": wk["d2h"],
                    "ms_per_step": e2e_ms_step, "api": wk["e2e_api"]},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "k_srcnn_tc (fused conv1+conv2+conv3, row-walking tcgen05)" if args.variant != "fp32" else "k_conv99x11_strict+k_conv55_strict",
                         "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"],
                         "frac_of_sustained_peak": achieved / peaks["bf16_sustained"], "peak_sustained": peaks["bf16_sustained"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
                         # of this command (the Y' plane the kernel writes stays in L2 for the merge kernel)
                         "traffic": traffic, "traffic_source": traffic_file,
                         "peak_source": peaks["src"], "kernel_ms": k_ms, "kernel_launches_per_step": launches_kb,
                         "algorithmic_flop_per_step": FLOP_PER_PX * px_rank},
            "stages": {"colour_bicubic_ms": a_ms / steps, "srcnn_ms": k_ms, "merge_ms": c_ms / steps,
                       "colour_bicubic_GBs": a_gbs, "colour_bicubic_frac_hbm": a_gbs / peaks["hbm"],
                       "merge_GBs": c_gbs, "merge_frac_hbm": c_gbs / peaks["hbm"], "hbm_peak_GBs": peaks["hbm"],
                       "merge": "fused into the SRCNN kernel's last epilogue" if fused_merge else "separate launch",
                       "timed_how": "srcnn_ms: CUDA events around every fused-kernel launch of the timed region; colour_bicubic_ms / merge_ms: "
                                    "a serialised pass of %d step(s) after it (events around every stage: no overlap between calls)" % ser_steps,
                       "serialised_ms_per_step": ser_ms_step,
                       # in the timed region the merge of step i and the colour+bicubic of step i+1 run side by side
                       # (programmatic dependent launch, two plane sets): time between consecutive fused-kernel launches
                       "between_srcnn_launches_ms": between_ms / (calls - 1) if (args.config == "cfg2" and calls > 1) else None,
                       "path_frac_of_tensor_peak": FLOP_PER_PX * px_rank / (ms_step * 1e-3) / 1e12 / peaks["bf16"]},
            "clocks": sampler.result(),
            "timed_region_s": timed_region_s,
        }
        if wk["e2e_note"]:
            line["e2e"]["note"] = wk["e2e_note"]
        if copy_ms_step == copy_ms_step:   # not NaN
            line["e2e"].update({"copy_only_ms_per_step": copy_ms_step,
                                "pcie_ceiling_GBs": world * (wk["h2d"] + wk["d2h"]) / (copy_ms_step * 1e-3) / 1e9,
                                "frac_of_copy_ceiling": copy_ms_step / e2e_ms_step,
                                "ceiling": "the step's H2D and D2H bytes alone, pinned, concurrently on two streams, all ranks at once (max over ranks)"})
        if world == 1 and not args.no_cpu:
            run, kind, threads, desc = cpu_reference_runner(s)
            v, sample, _ = time_cpu(run, cfg, args.cpu_budget)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "what": desc}
            if kind == "reference":   # also as the reference's Makefile builds it
                line["cpu_baseline"]["makefile_flags"] = time_cpu_makefile_flags(cfg)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# the same configurations through the in-library multi-GPU driver: ONE process, srcnn_mgpu_* fans the step out
# ------------------------------------------------------------------------------------------------
def run_mgpu(args, cfg):
    import torch
    import srcnn_cpp_b200 as S
    n = args.gpus
    m = S.MultiEngine(list(range(n)), S.VARIANT_FP32 if args.variant == "fp32" else S.VARIANT_TC)
    W, H, s = cfg["w"], cfg["h"], cfg["scale"]
    OW, OH = S.out_dims(W, H, s)
    devs = [torch.device("cuda", d) for d in m.devices]
    if args.config == "cfg4":
        plan = m.band_plan(H, s)
        srcs = [synth_device(torch, devs[i], 99, (plan[i][3] - plan[i][2], W, 3)) for i in range(n)]
        dsts = [torch.empty((plan[i][1] - plan[i][0], OW, 3), dtype=torch.uint8, device=devs[i]) for i in range(n)]
        px_step = OW * OH

        def step():
            m.process_banded_device(srcs, W, H, s, dsts)
    else:
        N = cfg["frames"] if args.config != "cfg2" else n
        counts = [len(range(i, N, n)) for i in range(n)]
        ring = 4 if args.config == "cfg5" else None
        srcs = [synth_device(torch, devs[i], 77 + i, (counts[i], H, W, 3)) for i in range(n)]
        nout = [c if ring is None else min(ring, c) for c in counts]
        dsts = [torch.empty((nout[i], OH, OW, 3), dtype=torch.uint8, device=devs[i]) for i in range(n)]
        px_step = N * OW * OH

        def step():
            if ring is None:
                m.process_batch_device(srcs, s, dsts)
                return [0.0] * n
            acc = [0.0] * n
            for f0 in range(0, max(counts), ring):
                m.process_batch_device([srcs[i][f0:f0 + ring] if f0 < counts[i] else None for i in range(n)], s,
                                       [dsts[i][:max(0, min(ring, counts[i] - f0))] if f0 < counts[i] else None for i in range(n)])
                ms, _ = m.last_timing()
                acc = [a + b for a, b in zip(acc, ms)]
            return acc
    for d in devs:
        torch.cuda.synchronize(d)
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = m.launches()
    dev_ms, wall_ms = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        acc = step()
        wall_ms.append((time.perf_counter() - t0) * 1e3)
        ms, _ = m.last_timing()
        dev_ms.append(max(acc) if (acc and max(acc) > 0) else max(ms))   # max over devices of the device-side time of the step
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = m.launches() - launches0
    ms_step = float(np.mean(dev_ms))
    line = {"metric": cfg["metric"], "value": px_step / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": cfg["scaling"] if args.config != "cfg2" else "weak",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate (tcgen05)" if args.variant != "fp32" else "f32 (strict)",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "variant": args.variant,
                       "parallelism": "ONE process, srcnn_mgpu_* over %d devices (one host thread + context per device), no collective" % n,
                       "timing": "max over devices of CUDA-event time around each device's share; wall_ms_per_step = host clock around the call"},
            "wall_ms_per_step": float(np.mean(wall_ms)), "value_wall": px_step / (float(np.mean(wall_ms)) * 1e-3) / 1e6,
            "gpu_launches": launches, "clocks": sampler.result(), "e2e": None, "cpu_baseline": None}
    print(json.dumps(line), flush=True)
    m.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--mgpu", action="store_true", help="one process, the library's multi-GPU driver (srcnn_mgpu_*) instead of one rank per GPU")
    ap.add_argument("--sustain", type=float, default=0.0, help="make the timed region at least this many seconds long")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    heavy = args.config != "cfg2"             # ~100 ms per step on one GPU
    if args.steps is None:
        args.steps = 5 if heavy else 50
    if args.warmup is None:
        args.warmup = 3 if heavy else 5
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core it can (set before libgomp loads)
            os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        run_reference_arm(args, cfg, rank, world)
        return
    if args.mgpu:
        run_mgpu(args, cfg)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it so that one process drives each GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
