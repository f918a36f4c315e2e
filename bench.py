#!/usr/bin/env python
"""bench.py -- headline benchmark of the SRCNN hot path (BASELINE.json: output MPix/s, 1080p -> 4K x2).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA library through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation

One "step" = one pass of the whole hot path (colour+bicubic -> fused SRCNN -> merge+colour back) over
one synthetic 1920x1080 BGR frame per rank, producing a 3840x2160 frame (configs[1]).  With N > 1 each
rank processes its own frames (frame sharding, SURVEY 8e: no collective on the data path) -> weak
scaling; `value` = all ranks' output pixels / max-over-ranks device time.

  value      : inputs resident in HBM, CUDA-event timed on the launching stream, K steps back to back,
               rotating over more frame/result buffers than fit in L2 (so no step reads a warm input)
  e2e        : the same metric through srcnn_process_host with pinned HOST buffers: H2D of the frame and
               D2H of the result are inside the timed region of every step
  roofline   : fused SRCNN kernel, 16 064 algorithmic FLOP per output pixel / its CUDA-event duration,
               against MEASURED_PEAKS.json's dense bf16 burst figure (same tensor rate as fp16)
  cpu_baseline: the reference's CPU code (oracle/_ref/libref.so + cv2, IPP off) on the box's host
               cores, on a bounded sample of the same workload
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, SCALE = 1920, 1080, 2.0
OW, OH = 3840, 2160
FLOP_PER_PX = 16064           # conv1 10368 + conv2 4096 + conv3 1600
METRIC = "output MPix/s (1080p->4K x2)"
UNIT = "MPix/s"
WORKLOAD = "single 1920x1080 -> 3840x2160 x2 synthetic BGR frame per rank per step (BASELINE configs[1])"


def synth_frame(seed, h=H, w=W):
    """Seeded synthetic BGR frame: smooth structure + grain (natural-like statistics, full 0..255 range)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.empty((h, w, 3), np.float32)
    for c in range(3):
        fx, fy, ph = rng.uniform(0.002, 0.02, 3)
        img[:, :, c] = 127 + 90 * np.sin(xx * fx * 6.28 + ph * 100) * np.cos(yy * fy * 6.28) + 30 * np.sin((xx + yy) * 0.05 * (c + 1))
    img += rng.normal(0, 12, img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=float(d["bf16_tflops"]), hbm=float(d["hbm_gbs"]), src="measured (MEASURED_PEAKS.json, burst)")
    return dict(bf16=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

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
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref when present, else the oracle port) -- the only place bench.py
# executes anything under oracle/.
# ------------------------------------------------------------------------------------------------
def cpu_reference_runner():
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
            return cv2_pipeline(img, SCALE, ref.cnn)
        desc = "reference src/srcnn.cpp conv functions compiled unmodified (-O3, bit-identical to the Makefile's -O0) + cv2 (IPP off) stages"
    except Exception:
        orc = Oracle()
        kind, threads = "port", orc.threads()

        def run(img):
            return orc.pipeline(img, SCALE)
        desc = "oracle/srcnn_oracle.c port (OpenMP)"
    return run, kind, threads, desc


def time_cpu(run, budget_s=12.0):
    """Times the CPU path on a bounded sample of the workload: a centred crop of the 1080p frame sized so
    one pass takes roughly budget_s (at most the whole frame).  Returns (MPix/s, sample description, seconds)."""
    frame = synth_frame(0)
    probe = np.ascontiguousarray(frame[:135, :240])
    run(probe)
    t = time.perf_counter(); run(probe); dt = time.perf_counter() - t
    rate = (270 * 480) / dt                       # output px/s on the probe
    want_px = min(OW * OH, max(270 * 480, rate * budget_s))
    frac = (want_px / (OW * OH)) ** 0.5
    h = max(135, min(H, int(H * frac) // 2 * 2)); w = max(240, min(W, int(W * frac) // 2 * 2))
    y0, x0 = (H - h) // 2, (W - w) // 2
    crop = np.ascontiguousarray(frame[y0:y0 + h, x0:x0 + w])
    t = time.perf_counter(); out = run(crop); dt = time.perf_counter() - t
    px = out.shape[0] * out.shape[1]
    return px / dt / 1e6, "%dx%d crop of the 1080p frame -> %dx%d (%.2f MPix out, %.1f s)" % (w, h, out.shape[1], out.shape[0], px / 1e6, dt), dt


def time_cpu_makefile_flags(budget_s=4.0):
    """The same CPU code as the reference's own Makefile really builds it (no -O flag on the compile lines, Makefile:21-23,43):
    oracle/_ref/libref_O0.so, bit-identical output, ~10x slower.  A small bounded sample; None when that build is absent."""
    try:
        from oracle.oracle import RefLib, cv2_pipeline
        ref = RefLib("O0")
    except Exception:
        return None

    def run(img):
        return cv2_pipeline(img, SCALE, ref.cnn)
    v, sample, _ = time_cpu(run, budget_s)
    return {"value": v, "unit": UNIT, "cores": ref.threads(), "sample": sample,
            "what": "same sources with the reference Makefile's own compile flags (-mtune=native -fopenmp, no -O)"}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    run, kind, threads, desc = cpu_reference_runner()
    run(np.ascontiguousarray(synth_frame(0)[:64, :64]))
    per_step_budget = max(1.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    vals, sample = [], ""
    for i in range(args.warmup + args.steps):
        v, sample, dt = time_cpu(run, per_step_budget)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals]) * 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU reference arm: each step is a bounded sample of the workload; rank 0 only"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "what": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
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

    # more frame/result buffers than L2 can hold: NBUF * (6.2 MB in + 24.9 MB out) = 249 MB > 126 MB
    NBUF = 8
    frames_h = [synth_frame(1000 * rank + i) for i in range(2)]
    src = [torch.from_numpy(frames_h[i % 2]).to(dev) for i in range(NBUF)]
    dst = [torch.empty((OH, OW, 3), dtype=torch.uint8, device=dev) for _ in range(NBUF)]
    px_step = OW * OH

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        eng.process_device(src[i % NBUF], SCALE, dst[i % NBUF])

    for i in range(args.warmup):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.profile_enable(True)
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launches - launches0
    stage_ms, calls = eng.profile_read()
    eng.profile_enable(False)

    # ---- end to end through the host-buffer C-ABI call: pinned host memory, H2D + D2H inside ----
    pin_in = [S.PinnedBuffer(H * W * 3) for _ in range(2)]
    pin_out = [S.PinnedBuffer(OH * OW * 3) for _ in range(2)]
    for k in range(2):
        pin_in[k].array[:] = frames_h[k].reshape(-1)
    L = eng.L

    def e2e_step(i):
        rc = L.srcnn_process_host(eng.ctx, pin_in[i % 2].ptr, W, H, 3 * W, S.ORDER_BGR, SCALE, pin_out[i % 2].ptr, 3 * OW)
        if rc != 0:
            raise S.SrcnnError(rc, L.srcnn_last_error(eng.ctx).decode())
    e2e_steps = max(3, min(args.steps, 20))
    for i in range(3):
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
    e2e_ms = max(f0.elapsed_time(f1), e2e_ms_wall)   # the call blocks until dst is ready: wall >= device
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # sanity: the timed path produced a plausible frame (not all zeros) -- cheap guard against silent no-ops
    chk = int(dst[(args.warmup + args.steps - 1) % NBUF][::97, ::101].to(torch.int32).sum().item())
    assert chk > 0

    # max over ranks
    t = torch.tensor([ms_total, e2e_ms / e2e_steps, stage_ms[0], stage_ms[1], stage_ms[2]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_step, a_ms, b_ms, c_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        peaks = load_peaks()
        ms_step = ms_total / args.steps
        value = world * px_step / (ms_step * 1e-3) / 1e6
        e2e_value = world * px_step / (e2e_ms_step * 1e-3) / 1e6
        k_ms = b_ms / max(1, calls)                       # fused SRCNN kernel, average launch duration
        achieved = FLOP_PER_PX * px_step / (k_ms * 1e-3) / 1e12
        a_bytes = (3.0 / (SCALE * SCALE) + 3.0) * px_step  # colour+bicubic: 3/s^2 read + 3 written per output px
        a_gbs = a_bytes / ((a_ms / max(1, calls)) * 1e-3) / 1e9
        fused_merge = c_ms / max(1, calls) < 5e-3     # merge + colour-back runs inside the fused kernel: no K-C launch
        c_gbs = 0.0 if fused_merge else 6.0 * px_step / ((c_ms / max(1, calls)) * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (strict)" if args.variant == "fp32" else "f16 operands, f32 accumulate (tcgen05)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "variant": args.variant, "parallelism": "frame-sharded x%d, no collective" % world,
                       "l2": "rotating %d frame/result buffers (%.0f MB) > 126 MB L2; no flush kernel in the timed region" % (NBUF, NBUF * (H * W * 3 + OH * OW * 3) / 1e6)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": H * W * 3, "d2h_bytes_per_step": OH * OW * 3,
                    "ms_per_step": e2e_ms_step, "api": "srcnn_process_host (pinned host buffers, blocking)"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "k_srcnn_tc2 (fused conv1+conv2+conv3, row-walking tcgen05)" if args.variant != "fp32" else "k_conv99x11_strict+k_conv55_strict",
                         "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this
                         # command (profiles/r1_summary.md); the Y' plane it writes stays in L2 for the merge kernel
                         "traffic": (8.51e6 if args.variant != "fp32" else None),
                         "peak_source": peaks["src"], "kernel_ms": k_ms,
                         "algorithmic_flop_per_launch": FLOP_PER_PX * px_step},
            "stages": {"colour_bicubic_ms": a_ms / max(1, calls), "srcnn_ms": k_ms, "merge_ms": c_ms / max(1, calls),
                       "colour_bicubic_GBs": a_gbs, "colour_bicubic_frac_hbm": a_gbs / peaks["hbm"],
                       "merge_GBs": c_gbs, "merge_frac_hbm": c_gbs / peaks["hbm"], "hbm_peak_GBs": peaks["hbm"],
                       "merge": "fused into the SRCNN kernel's last epilogue" if fused_merge else "separate launch"},
            "clocks": sampler.result(),
        }
        if world == 1 and not args.no_cpu:
            run, kind, threads, desc = cpu_reference_runner()
            v, sample, _ = time_cpu(run, args.cpu_budget)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "what": desc}
            if kind == "reference":   # also as the reference's Makefile builds it
                line["cpu_baseline"]["makefile_flags"] = time_cpu_makefile_flags()
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core it can (set before libgomp loads)
            os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        run_reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it so that one process drives each GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
