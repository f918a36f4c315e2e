"""The integer-scale colour+bicubic kernel (csrc/color_bicubic_int.cu: x2 / x4, constant tap phases, one warp per plane walking
down the footprint rows) against the oracle, bit for bit, on every path it has: 32-bit loads (4-byte aligned rows), byte loads
(anything else), image-edge tiles, one tile and several tiles per axis, BGR and RGB order, 5-, 8- and 16-row tiles, bands that begin
and end on odd output rows, and a batch of frames in one launch.
Replaces cvtColor(BGR2YCrCb) src/srcnn.cpp:509, split :540 and resize(..., CV_INTER_CUBIC) :570-583."""
import os

import numpy as np
import pytest

from conftest import diff_stats

pytestmark = pytest.mark.gpu


def _planes(oh, ow):
    import torch
    pitch = (ow + 127) // 128 * 128
    return [torch.zeros((oh, pitch), dtype=torch.uint8, device="cuda:0")[:, :ow] for _ in range(3)]


def _want(oracle, img_bgr, ow, oh):
    ycc = oracle.bgr2ycrcb(img_bgr)
    return [oracle.resize_cubic(np.ascontiguousarray(ycc[:, :, k]), ow, oh) for k in range(3)]


def _check(engine, oracle, src_dev, img_bgr, scale, order):
    import srcnn_cpp_b200 as S
    h, w, _ = img_bgr.shape
    ow, oh = S.out_dims(w, h, scale)
    pl = _planes(oh, ow)
    engine.stage_color_bicubic(src_dev, scale, *pl, order=order)
    engine.sync()
    for k, want in enumerate(_want(oracle, img_bgr, ow, oh)):
        got = pl[k].cpu().numpy()
        assert np.array_equal(got, want), (k, diff_stats(got, want))


# (w, h): widths that are / are not multiples of 16 and of 4, one and several 256-column tiles, heights around the 8- and 16-row
# tile sizes and the 3-row apron
GEOMS = [(16, 1), (16, 2), (32, 3), (48, 9), (20, 17), (132, 40), (272, 33), (260, 19), (12, 5), (400, 70), (640, 9)]


@pytest.mark.parametrize("scale", [2.0, 4.0])
@pytest.mark.parametrize("w,h", GEOMS)
def test_integer_scales_bit_exact(engine, oracle, w, h, scale):
    import srcnn_cpp_b200 as S
    import torch
    rng = np.random.default_rng(w * 131 + h * 7 + int(scale))
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    _check(engine, oracle, torch.from_numpy(img).to("cuda:0"), img, scale, S.ORDER_BGR)
    _check(engine, oracle, torch.from_numpy(np.ascontiguousarray(img[:, :, ::-1])).to("cuda:0"), img, scale, S.ORDER_RGB)


@pytest.mark.parametrize("shift", [1, 2, 4, 8])
def test_unaligned_source_rows(engine, oracle, shift):
    """source pointer / stride that keep (shift 4, 8) or rule out (shift 1, 2) the 32-bit loads -- same bytes out"""
    import srcnn_cpp_b200 as S
    import torch
    w, h = 64, 21
    rng = np.random.default_rng(shift)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    stride = w * 3 + 16 + shift
    buf = torch.zeros(h * stride + 64, dtype=torch.uint8, device="cuda:0")
    view = buf[shift:shift + h * stride].view(h, stride)[:, :w * 3].view(h, w, 3)   # rows `stride` apart, first byte at `shift`
    view.copy_(torch.from_numpy(img).to("cuda:0"))
    assert view.stride(0) == stride and view.data_ptr() % 16 == shift % 16
    _check(engine, oracle, view, img, 2.0, S.ORDER_BGR)


def test_extreme_pixel_values(engine, oracle):
    """saturating chroma, negative horizontal sums (the int -> float trick of the kernel must hold for both signs)"""
    import srcnn_cpp_b200 as S
    import torch
    w, h = 96, 24
    rng = np.random.default_rng(3)
    img = (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)      # every channel 0 or 255
    img[:, ::2] = img[:, ::2][:, :, ::-1]
    _check(engine, oracle, torch.from_numpy(img).to("cuda:0"), img, 2.0, S.ORDER_BGR)
    _check(engine, oracle, torch.from_numpy(img).to("cuda:0"), img, 4.0, S.ORDER_BGR)


@pytest.mark.parametrize("scale,r0,r1", [(2.0, 0, 1), (2.0, 1, 2), (2.0, 3, 38), (2.0, 17, 80), (4.0, 1, 6), (4.0, 7, 91), (4.0, 155, 160)])
def test_bands_on_odd_rows_equal_the_whole_image(engine, scale, r0, r1):
    """the tile grid of a band starts at the band's first row: odd first rows, one-row bands, the image's last rows"""
    import srcnn_cpp_b200 as S
    import torch
    w, h = 80, 40
    rng = np.random.default_rng(int(scale) * 100 + r0)
    img = torch.from_numpy(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).to("cuda:0")
    ow, oh = S.out_dims(w, h, scale)
    whole = torch.zeros((oh, ow, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_device(img, scale, whole)
    s0, s1 = S.band_src_rows(h, scale, r0, r1)
    band = torch.zeros((r1 - r0, ow, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_band_device(img[s0:s1].contiguous(), w, h, s0, s1, scale, r0, r1, band)
    engine.sync()
    assert torch.equal(band, whole[r0:r1])


def test_generic_and_integer_scale_kernels_agree():
    """SRCNN_KA_INT=0 keeps the generic tiled kernel for x2: both must produce the same planes, incl. a batch in one launch
    and several tile heights of the integer-scale kernel"""
    import srcnn_cpp_b200 as S
    import torch
    w, h, n = 272, 150, 3
    rng = np.random.default_rng(11)
    frames = torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).to("cuda:0")
    ow, oh = S.out_dims(w, h, 2.0)
    outs = {}
    for name, env in (("generic", {"SRCNN_KA_INT": "0"}), ("int8", {"SRCNN_KA_ISR": "8"}), ("int16", {"SRCNN_KA_ISR": "16"}),
                      ("int5", {"SRCNN_KA_ISR": "5"})):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            eng = S.Engine(device=0, variant=S.VARIANT_TC)     # the knobs are read when a context is created
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        dst = torch.zeros((n, oh, ow, 3), dtype=torch.uint8, device="cuda:0")
        torch.cuda.synchronize()
        eng.process_batch_device(frames, 2.0, dst)
        eng.sync()
        outs[name] = dst.cpu().numpy()
        eng.close()
    for name in ("int8", "int16", "int5"):
        assert np.array_equal(outs[name], outs["generic"]), name
