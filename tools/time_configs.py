"""Device-resident throughput of the other BASELINE configurations on one B200 (not bench lines: bench.py measures configs[1]).
CUDA events on the stream the kernels are launched on, 3 warm-up + 5 timed calls each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srcnn_cpp_b200 as S

st = torch.cuda.Stream()
torch.cuda.set_stream(st)
eng = S.Engine(0, stream=st.cuda_stream)


def timed(fn, n=5, warm=3):
    for _ in range(warm):
        fn()
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    eng.sync()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# configs[2]: 1024 x (1280x720 -> 2560x1440), one batch call
src = torch.randint(0, 256, (1024, 720, 1280, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty((1024, 1440, 2560, 3), dtype=torch.uint8, device="cuda")
ms = timed(lambda: eng.process_batch_device(src, 2.0, dst), n=3, warm=1)
print("cfg3 1024 x 720p -> 1440p: %.2f ms per batch, %.1f GPix/s" % (ms, 1024 * 1440 * 2560 / ms / 1e6))
del src, dst
torch.cuda.empty_cache()
# configs[4]: 3840x2160 -> 15360x8640 x4
src = torch.randint(0, 256, (2160, 3840, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty((8640, 15360, 3), dtype=torch.uint8, device="cuda")
ms = timed(lambda: eng.process_device(src, 4.0, dst))
print("cfg5 4K -> 16K x4: %.3f ms per frame, %.1f GPix/s" % (ms, 8640 * 15360 / ms / 1e6))
del src, dst
torch.cuda.empty_cache()
# configs[3]: 32768^2 -> 65536^2
src = torch.randint(0, 256, (32768, 32768, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty((65536, 65536, 3), dtype=torch.uint8, device="cuda")
ms = timed(lambda: eng.process_device(src, 2.0, dst), n=2, warm=1)
print("cfg4 32768^2 -> 65536^2: %.1f ms per image, %.1f GPix/s" % (ms, 65536 * 65536 / ms / 1e6))
