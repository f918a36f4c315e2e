"""Debug: per-role clock64 timeline of pipeline 0 / CTA 0 of the row-walking fused kernel (SRCNN_TC_DEBUG=1)."""
import ctypes as C, os, sys
os.environ["SRCNN_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import srcnn_cpp_b200 as S
st = torch.cuda.Stream()
torch.cuda.set_stream(st)   # torch and the context on one stream
eng = S.Engine(0, stream=st.cuda_stream)
y = torch.randint(0, 256, (2160, 3840), dtype=torch.uint8, device="cuda")
out = torch.zeros_like(y)
for _ in range(3):
    eng.stage_cnn(y, out)
eng.sync()
buf = (C.c_longlong * (4 * 64 * 8))()
eng.L.srcnn_debug_tc2_timeline.argtypes = [C.c_void_p, C.c_void_p]
assert eng.L.srcnn_debug_tc2_timeline(eng.ctx, buf) == 0
a = np.array(buf[:]).reshape(4, 64, 8)
t0 = a[a > 0].min()
lo, hi = 30, 40
print("P (ring row t; conv1 of row t-8): fetch+gather | st+stage | wait::st | ringfree wait | named bar | unitfree wait | issue conv1")
for i in range(lo, hi):
    r = a[1, i]
    print(" t%2d  t=%7d  gather %5d  st %5d  waitst %5d  freew %5d  bar %5d  unitw %5d  issue %5d   period %5d" % (
        i, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[7] - r[6], a[1, i + 1, 0] - r[0]))
print("E1 (row i): D1full wait | work + arrive | wait all arrived | issue conv2")
for i in range(lo, hi):
    r = a[0, i]
    print(" i%2d  t=%7d  wait %5d  work %5d  allwait %5d  issue %5d   period %5d" % (i, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], a[0, i + 1, 0] - r[0]))
print("E2 (row i): D2full wait | work + arrive | wait all arrived | issue conv3")
for i in range(lo, hi):
    r = a[3, i]
    print(" i%2d  t=%7d  wait %5d  work %5d  allwait %5d  issue %5d   period %5d" % (i, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], a[3, i + 1, 0] - r[0]))
print("E3 (row): Tfull wait | T load + freed | tap adds + publish + bar (row-2) | exchange read, store")
for i in range(lo, hi):
    r = a[2, i]
    print(" i%2d  t=%7d  wait %5d  load %5d  adds+bar %5d  out %5d   step %5d   period %5d" % (
        i, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[4] - r[0], a[2, i + 1, 0] - r[0]))
i = 34
ev = [("conv1 issue", a[1, i + 8, 6]), ("conv1 issued", a[1, i + 8, 7]), ("D1full seen", a[0, i, 1]), ("A1ready", a[0, i, 2]), ("conv2 issue", a[0, i, 3]), ("conv2 issued", a[0, i, 4]),
      ("D2 seen", a[3, i, 1]), ("A2ready", a[3, i, 2]), ("conv3 issue", a[3, i, 3]), ("conv3 issued", a[3, i, 4]), ("T seen", a[2, i, 1]), ("T loaded, unit freed", a[2, i, 2]), ("E3 step done", a[2, i, 4])]
print("row %d life: " % i + "  ".join("%s +%d" % (n, v - ev[0][1]) for n, v in ev))
