"""frawscale-compatible resize stage: bit-identical to the reference's FRAWResizeEngine::scale
-- against vectors generated from the compiled reference (tools/make_fraw_golden.py) and, where
oracle/_ref/libfraw.so is present, against the compiled reference itself on fresh inputs."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "fraw_golden.npz"))


def test_golden_vectors_match_compiled_reference():
    """CPU: the committed vectors still equal what the compiled reference produces (where it is available)."""
    from oracle.oracle import FrawRef
    if not FrawRef.available():
        pytest.skip("oracle/_ref/libfraw.so not built and /root/reference absent")
    ref, g = FrawRef(), _golden()
    k = 0
    while "cfg%d" % k in g:
        sw, sh, dw, dh, flt = g["cfg%d" % k].tolist()
        got = ref.scale(g["src%d" % k], dw, dh, flt)
        assert np.array_equal(got.view(np.uint32), g["dst%d" % k].view(np.uint32)), k
        k += 1
    assert k >= 10


@pytest.mark.gpu
def test_gpu_matches_golden_vectors(engine):
    import torch
    g = _golden()
    k = 0
    while "cfg%d" % k in g:
        sw, sh, dw, dh, flt = g["cfg%d" % k].tolist()
        src = torch.from_numpy(g["src%d" % k]).to("cuda:0")
        dst = torch.zeros((dh, dw), dtype=torch.float32, device="cuda:0")
        engine.fraw_scale(src, dst, flt)
        engine.sync()
        assert np.array_equal(dst.cpu().numpy().view(np.uint32), g["dst%d" % k].view(np.uint32)), (k, sw, sh, dw, dh, flt)
        k += 1


@pytest.mark.gpu
@pytest.mark.parametrize("sw,sh,dw,dh,flt", [(320, 180, 640, 360, 2), (123, 77, 185, 115, 2), (200, 100, 100, 50, 2),
                                             (64, 48, 256, 192, 1), (64, 48, 256, 192, 0), (50, 40, 75, 20, 2)])
def test_gpu_matches_compiled_reference(engine, sw, sh, dw, dh, flt):
    import torch
    from oracle.oracle import FrawRef
    if not FrawRef.available():
        pytest.skip("oracle/_ref/libfraw.so did not travel")
    rng = np.random.default_rng(sw * 7 + dh)
    src = (rng.random((sh, sw), dtype=np.float32) * 255).astype(np.float32)
    want = FrawRef().scale(src, dw, dh, flt)
    dst = torch.zeros((dh, dw), dtype=torch.float32, device="cuda:0")
    engine.fraw_scale(torch.from_numpy(src).to("cuda:0"), dst, flt)
    engine.sync()
    assert np.array_equal(dst.cpu().numpy().view(np.uint32), want.view(np.uint32))
