"""Parity anchors the round-1 review asked for ("Parity hardening"), all on the GPU box:
  * the tcgen05 and FP32 kernels against the COMPILED REFERENCE itself (oracle/_ref/libref.so = the unmodified
    src/srcnn.cpp, Convolution99x11 :254-325 + Convolution55 :189-243), on the same machine in the same process;
  * configs[4] (x4, 3840x2160 -> 15360x8640) and configs[2] (1280x720 -> 2560x1440) at FULL size with real content:
    256x320 crops (four corners = every border case, plus the centre; at x4 also crops straddling strip cuts) of the
    whole-path BGR result against the oracle.
Tolerances (north_star): Y' of the tensor-core path <= 1 LSB on >= 99.9 % of pixels, max 2; FP32 variant and the
colour/bicubic/merge stages bit-exact.  On BGR bytes the bound is DERIVED: |dB| <= |dY'| and the same for G and R before
saturation (the inverse colour transform adds Y' to a chroma term that is identical on both sides, src/srcnn.cpp:657 /
SURVEY A.1), so BGR inherits max 2; the tests assert that, and report the observed maximum."""
import numpy as np
import pytest

from conftest import diff_stats

pytestmark = pytest.mark.gpu

TC_MAX, TC_LE1 = 2, 0.999


def _synth(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.empty((h, w, 3), np.float32)
    for c in range(3):
        img[:, :, c] = 127 + 80 * np.sin(xx * 0.031 * (c + 1)) * np.cos(yy * 0.023) + 25 * np.sin((xx + yy) * 0.11)
    img += rng.normal(0, 10, img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w", [(96, 140), (33, 250), (260, 131)])
def test_kernels_against_the_compiled_reference(engine, reflib, h, w):
    """GPU kernels vs RefLib().cnn -- the reference's own object code, not our restatement of it."""
    import torch
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(h * w)
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    want, act2 = reflib.cnn(y, want_act2=True)
    d_y = torch.from_numpy(y).to("cuda:0")
    out = torch.zeros((h, w), dtype=torch.uint8, device="cuda:0")
    engine.stage_cnn(d_y, out, variant=S.VARIANT_FP32)
    engine.sync()
    assert np.array_equal(out.cpu().numpy(), want)                      # strict FP32: the reference's bytes
    d_act2 = torch.zeros((32, h, w), dtype=torch.float32, device="cuda:0")
    engine.stage_conv99x11_fp32(d_y, d_act2)
    engine.sync()
    assert np.array_equal(d_act2.cpu().numpy().view(np.uint32), np.asarray(act2).reshape(32, h, w).view(np.uint32))   # act2 bit patterns
    out.zero_()
    engine.stage_cnn(d_y, out, variant=S.VARIANT_TC)
    engine.sync()
    st = diff_stats(out.cpu().numpy(), want)
    assert st["max"] <= TC_MAX and st["le1"] >= TC_LE1, st


def _crop_check(engine, oracle, img, scale, crops, ch=256, cw=320):
    """Whole-path BGR of `img` on the GPU; each crop compared with the oracle: planes resized whole (cheap), the CNN on
    the crop with 6 px of context where the image continues.  Returns per-crop stats."""
    import torch
    import srcnn_cpp_b200 as S
    h, w, _ = img.shape
    ow, oh = S.out_dims(w, h, scale)
    d = torch.from_numpy(img).to("cuda:0")
    out = torch.zeros((oh, ow, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_device(d, scale, out)
    engine.sync()
    ycc = oracle.bgr2ycrcb(img)
    up = [oracle.resize_cubic(np.ascontiguousarray(ycc[:, :, k]), ow, oh) for k in range(3)]
    m = 6
    stats = []
    for (r0, c0) in crops(oh, ow, ch, cw):
        ra, rb, ca, cb = max(r0 - m, 0), min(r0 + ch + m, oh), max(c0 - m, 0), min(c0 + cw + m, ow)
        yref = oracle.cnn(np.ascontiguousarray(up[0][ra:rb, ca:cb]))[r0 - ra:r0 - ra + ch, c0 - ca:c0 - ca + cw]
        ref = oracle.ycrcb2bgr(np.ascontiguousarray(np.dstack([yref, up[1][r0:r0 + ch, c0:c0 + cw], up[2][r0:r0 + ch, c0:c0 + cw]])))
        got = out[r0:r0 + ch, c0:c0 + cw].cpu().numpy()
        st = diff_stats(got, ref)
        stats.append(((r0, c0), st))
        assert st["max"] <= TC_MAX and st["le1"] >= TC_LE1, ((r0, c0), st)
    del out, d
    return stats


def _corners_and_centre(oh, ow, ch, cw):
    return [(0, 0), (0, ow - cw), (oh - ch, 0), (oh - ch, ow - cw), ((oh - ch) // 2, (ow - cw) // 2)]


def test_cfg5_x4_full_size_crops_vs_oracle(engine, oracle):
    """configs[4]: one real-content 3840x2160 -> 15360x8640 frame; corners, centre, and two crops that straddle strip
    boundaries (columns 124 k) and a pipeline's segment cut far from any border."""
    rng = np.random.default_rng(55)
    img = _synth(rng, 2160, 3840)

    def crops(oh, ow, ch, cw):
        return _corners_and_centre(oh, ow, ch, cw) + [(4000, 124 * 60 - 160), (7000, 124 * 100 - 10)]
    stats = _crop_check(engine, oracle, img, 4.0, crops)
    assert len(stats) == 7


def test_cfg3_720p_frame_corners_vs_oracle(engine, oracle):
    """configs[2]: one real-content 1280x720 -> 2560x1440 frame, corners + centre against the oracle."""
    rng = np.random.default_rng(56)
    img = _synth(rng, 720, 1280)
    _crop_check(engine, oracle, img, 2.0, _corners_and_centre)


def test_bgr_bound_is_inherited_from_y(engine, oracle):
    """The derived BGR bound: with identical Cr/Cb on both sides, max|dBGR| <= max|dY'| (saturation can only shrink it)."""
    import torch
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(57)
    img = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)      # uniform noise: the worst case, 12-18 % of outputs saturate
    want_bgr, st = oracle.pipeline(img, 2.0, stages=True)
    want_y = st["cnn_y"]
    d = torch.from_numpy(img).to("cuda:0")
    y, cr, cb = [torch.zeros((400, 640), dtype=torch.uint8, device="cuda:0")[:, :600] for _ in range(3)]
    engine.stage_color_bicubic(d, 2.0, y, cr, cb)
    yo = torch.zeros((400, 640), dtype=torch.uint8, device="cuda:0")[:, :600]   # same pitch as the other planes
    engine.stage_cnn(y, yo, variant=S.VARIANT_TC)
    out = torch.zeros((400, 600, 3), dtype=torch.uint8, device="cuda:0")
    engine.stage_merge(yo, cr, cb, out)
    engine.sync()
    dy = np.abs(yo.cpu().numpy().astype(np.int16) - want_y.astype(np.int16))
    db = np.abs(out.cpu().numpy().astype(np.int16) - want_bgr.astype(np.int16))
    assert dy.max() <= TC_MAX and (dy <= 1).mean() >= TC_LE1
    assert db.max() <= dy.max()
    assert (db.max(axis=2) <= dy).all()            # per pixel, every channel moves by at most what Y' moved
