"""Generates tests/golden/fraw_golden.npz from the reference's own frawscale.cpp (compiled unmodified into
oracle/_ref/libfraw.so).  Run in the authoring container (needs /root/reference); the vectors are committed so
that the GPU box can check the frawscale-compatible stage even if oracle/_ref did not travel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import FrawRef  # noqa: E402

CASES = [(24, 18, 48, 36, 2), (24, 18, 36, 27, 2), (31, 17, 93, 68, 2), (40, 30, 20, 15, 2), (33, 21, 50, 13, 2),
         (24, 18, 48, 36, 1), (24, 18, 48, 36, 0), (40, 30, 17, 45, 1), (7, 5, 28, 20, 2), (1, 1, 4, 4, 2)]


def main():
    ref = FrawRef()
    out = {}
    for k, (sw, sh, dw, dh, flt) in enumerate(CASES):
        rng = np.random.default_rng(100 + k)
        src = (rng.random((sh, sw), dtype=np.float32) * 255).astype(np.float32)
        out["src%d" % k] = src
        out["dst%d" % k] = ref.scale(src, dw, dh, flt)
        out["cfg%d" % k] = np.array([sw, sh, dw, dh, flt], np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fraw_golden.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
