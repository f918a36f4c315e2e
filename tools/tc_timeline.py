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
buf = (C.c_longlong * (2 * 24 * 8))()
eng.L.srcnn_debug_tc_timeline.argtypes = [C.c_void_p, C.c_void_p]
assert eng.L.srcnn_debug_tc_timeline(eng.ctx, buf) == 0
a = np.array(buf[:]).reshape(2, 24, 8)
names = ["top", "c1done", "E1done", "c2done", "E2done", "c3done", "E3a done"]
for pipe in range(2):
    print("pipe", pipe)
    for g in range(2, 12):
        r = a[pipe, g]
        prev = a[pipe, g - 1][6]
        pb = a[pipe, g - 1][7]
        print(" g%2d" % g, "bar %5d E3b %5d | fetch+c1wait %5d E1 %5d c2wait %5d E2 %5d c3wait %5d E3a %5d | total %5d" % (
            pb - prev, r[0] - pb, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[6] - prev))
