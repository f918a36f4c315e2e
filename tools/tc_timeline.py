"""Debug: per-phase clock64 timeline of CTA 0 of the fused kernel (SRCNN_TC_DEBUG=1)."""
import ctypes as C, os, sys
os.environ["SRCNN_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import srcnn_cpp_b200 as S
eng = S.Engine(0)
y = torch.randint(0, 256, (2160, 3840), dtype=torch.uint8, device="cuda")
out = torch.zeros_like(y)
for _ in range(3):
    eng.stage_cnn(y, out)
eng.sync()
buf = (C.c_longlong * (2 * 24 * 16))()
eng.L.srcnn_debug_tc_timeline.argtypes = [C.c_void_p, C.c_void_p]
assert eng.L.srcnn_debug_tc_timeline(eng.ctx, buf) == 0
a = np.array(buf[:]).reshape(2, 24, 16)
names = ["c1a", "E1a", "c1b", "E1b", "c2a", "E2a", "c2b", "E2b", "c3a", "E3a", "c3b", "E3b'", "bar", "E3b"]
for pipe in range(2):
    print("pipe", pipe)
    for g in range(2, 10):
        r = a[pipe, g]
        print(" g%2d " % g + " ".join("%s %4d" % (names[k], r[k + 1] - r[k]) for k in range(14)) + " | total %5d" % (r[14] - r[0]))
