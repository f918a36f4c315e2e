import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/srcnn_oracle.c).  Test infrastructure only."""
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    """The reference's own conv code compiled unmodified (oracle/_ref/libref.so), if present."""
    from oracle.oracle import RefLib
    if not RefLib.available():
        pytest.skip("oracle/_ref/libref.so not built and /root/reference absent")
    return RefLib()


@pytest.fixture(scope="session")
def golden():
    import cv2
    g = os.path.join(ROOT, "tests", "golden")
    src = cv2.imread(os.path.join(g, "butterfly.png"))
    dst = cv2.imread(os.path.join(g, "butterfly-srcnn.png"))
    assert src is not None and dst is not None
    return src, dst


@pytest.fixture(scope="session")
def engine():
    """One srcnn_ctx on cuda:0 through the C ABI.  Fails loudly if the library or the GPU is missing."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import srcnn_cpp_b200 as S
    eng = S.Engine(device=0, variant=S.VARIANT_TC)
    # The context's own stream is non-blocking: it does not order itself against torch's default stream, so a test's
    # torch.zeros / copy_ could still be in flight when a kernel of ours reads or writes the same tensor (seen as an all-zero
    # result under compute-sanitizer, whose slower kernels widen the window).  One torch stream for both sides instead:
    # everything a test does on the device is then in stream order.
    st = torch.cuda.Stream(device=0)
    torch.cuda.synchronize()
    torch.cuda.set_stream(st)
    eng.set_stream(st.cuda_stream)
    yield eng
    eng.close()


def natural_like(rng, h, w):
    """Smooth-ish synthetic BGR image (low-res noise upsampled + gradients + a little grain)."""
    import cv2
    small = rng.integers(0, 256, (max(2, h // 8), max(2, w // 8), 3), dtype=np.uint8)
    img = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC).astype(np.int16)
    gx = np.linspace(0, 40, w, dtype=np.float32)[None, :, None]
    gy = np.linspace(0, 25, h, dtype=np.float32)[:, None, None]
    img = img + gx.astype(np.int16) + gy.astype(np.int16) + rng.integers(-6, 7, (h, w, 3), dtype=np.int16)
    return np.clip(img, 0, 255).astype(np.uint8)


def diff_stats(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return dict(max=int(d.max()), exact=float((d == 0).mean()), le1=float((d <= 1).mean()))
