"""Debug: error map of the fused tcgen05 kernel against the oracle on small noise images."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import srcnn_cpp_b200 as S
from oracle.oracle import Oracle
orc = Oracle()
eng = S.Engine(0)
shapes = [(40, 52), (20, 130), (300, 200), (1, 17), (17, 1)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (h, w) in shapes:
    rng = np.random.default_rng(h * 1000 + w + 1)
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    out = torch.zeros((h, w), dtype=torch.uint8, device="cuda:0")
    eng.stage_cnn(torch.from_numpy(y).cuda(), out, variant=S.VARIANT_TC)
    eng.sync()
    got, want = out.cpu().numpy().astype(int), orc.cnn(y).astype(int)
    d = np.abs(got - want)
    bad = d > 1
    print(f"== {h}x{w}: max {d.max()} exact {np.mean(d == 0):.4f} le1 {np.mean(d <= 1):.4f} bad {bad.sum()}")
    if bad.any():
        rows = np.where(bad.any(axis=1))[0]
        cols = np.where(bad.any(axis=0))[0]
        print("  bad rows:", rows[:40], "..." if len(rows) > 40 else "")
        print("  bad cols:", cols[:40], "..." if len(cols) > 40 else "")
        print("  bad per row (first 24 rows):", bad.sum(axis=1)[:24])
        r0 = rows[0]
        print(f"  row {r0} got :", got[r0, :24])
        print(f"  row {r0} want:", want[r0, :24])
