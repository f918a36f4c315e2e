"""CPU tests that PIN the oracle (oracle/srcnn_oracle.c) before anything is checked against it:
the reference's only golden vector, real OpenCV (python cv2, IPP off) stage by stage, and the
reference's own conv functions compiled unmodified (oracle/_ref/libref.so)."""
import hashlib

import numpy as np
import pytest

from conftest import natural_like


def test_blob_matches_reference_tables(reflib):
    from oracle.oracle import load_params
    assert np.array_equal(reflib.params().view(np.uint32), load_params().view(np.uint32))


def test_golden_butterfly_x1_5_bit_exact(oracle, golden):
    """Pictures/butterfly.png --scale=1.5 -> Pictures/butterfly-srcnn.png (reference README.md:34-45)."""
    src, dst = golden
    out, st = oracle.pipeline(src, 1.5, stages=True)
    assert out.shape == dst.shape == (576, 576, 3)
    assert np.array_equal(out, dst)
    # stage digests recorded in SURVEY.md Appendix B (computed there with cv2 + compiled reference)
    assert hashlib.sha256(st["up_y"].tobytes()).hexdigest() == "e4862936d82de0eba7d5f49f85e6e478bdeb6cd71fc074494788b660ba463be9"
    assert hashlib.sha256(st["cnn_y"].tobytes()).hexdigest() == "a9ab5bdb29e6086e6143fd32a7b2e2758b364e9fc8e54caa6fa5857b0cede232"
    assert hashlib.sha256(out.tobytes()).hexdigest() == "d49d417c868ce712d6d9ad8d10ad11e89e3a38791c6743d622afc1ece27ea12f"


def test_golden_via_cv2_and_compiled_reference(reflib, golden):
    from oracle.oracle import cv2_pipeline
    src, dst = golden
    assert np.array_equal(cv2_pipeline(src, 1.5, reflib.cnn), dst)


def test_colour_matches_cv2(oracle):
    import cv2
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (131, 257, 3), dtype=np.uint8)
    # include the extremes explicitly
    img[0, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [0, 255, 255], [255, 0, 255]]
    assert np.array_equal(oracle.bgr2ycrcb(img), cv2.cvtColor(img, cv2.COLOR_BGR2YCrCb))
    assert np.array_equal(oracle.ycrcb2bgr(img), cv2.cvtColor(img, cv2.COLOR_YCrCb2BGR))


@pytest.mark.parametrize("sw,sh,scale", [(96, 64, 2.0), (33, 17, 2.0), (50, 41, 3.0), (64, 48, 4.0), (37, 29, 1.5),
                                         (20, 13, 1.25), (7, 5, 2.0), (1, 1, 2.0), (3, 9, 2.7), (101, 77, 1.1),
                                         (64, 64, 0.5), (90, 70, 0.3)])
def test_resize_matches_cv2(oracle, sw, sh, scale):
    import cv2
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(sw * 1000 + sh)
    p = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    dw, dh = oracle.out_dims(sw, sh, scale)
    assert np.array_equal(oracle.resize_cubic(p, dw, dh), cv2.resize(p, (dw, dh), interpolation=cv2.INTER_CUBIC))


def test_x2_and_x4_tap_tables(oracle):
    """exact integer taps for the dyadic scales."""
    ofs, coef = oracle.cubic_taps(100, 200)
    assert coef[2].tolist() == [-72, 536, 1800, -216] and coef[3].tolist() == [-216, 1800, 536, -72]
    ofs4, coef4 = oracle.cubic_taps(100, 400)
    assert coef4[4].tolist() == [-135, 873, 1535, -225] and coef4[5].tolist() == [-21, 235, 1981, -147]


@pytest.mark.parametrize("h,w", [(40, 52), (9, 13), (1, 1), (3, 70), (64, 5)])
def test_cnn_matches_compiled_reference(oracle, reflib, h, w):
    rng = np.random.default_rng(h * 100 + w)
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    a, b = oracle.conv99x11(y), reflib.conv99x11(y)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(oracle.conv55(a), reflib.conv55(b))


def test_cnn_reference_O0_equals_O3(reflib):
    from oracle.oracle import RefLib
    r0 = RefLib("O0")
    rng = np.random.default_rng(7)
    y = rng.integers(0, 256, (24, 31), dtype=np.uint8)
    assert np.array_equal(r0.cnn(y), reflib.cnn(y))


def test_pipeline_matches_cv2_plus_reference_on_synthetic(oracle, reflib):
    from oracle.oracle import cv2_pipeline
    rng = np.random.default_rng(3)
    for (h, w, s) in [(31, 45, 2.0), (26, 22, 3.0), (40, 40, 1.5)]:
        img = natural_like(rng, h, w)
        assert np.array_equal(oracle.pipeline(img, s), cv2_pipeline(img, s, reflib.cnn))


def test_bad_ratio(oracle):
    img = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(ValueError):
        oracle.pipeline(img, 0.1)  # (int)(4*0.1) == 0 -> empty image, src/srcnn.cpp:485-495
