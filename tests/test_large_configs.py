"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the oracle needs ~1 us
per pixel, so whole-frame oracle runs are kept to one 1080p->4K frame, and the larger configs are checked on
crops, band/batch equivalence and invariants)."""
import numpy as np
import pytest

from conftest import diff_stats

pytestmark = pytest.mark.gpu


def _synth(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.empty((h, w, 3), np.float32)
    for c in range(3):
        img[:, :, c] = 127 + 80 * np.sin(xx * 0.031 * (c + 1)) * np.cos(yy * 0.023) + 25 * np.sin((xx + yy) * 0.11)
    img += rng.normal(0, 10, img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def test_cfg2_full_frame_vs_oracle(engine, oracle):
    """configs[1]: one 1920x1080 -> 3840x2160 frame.  K-A/K-C bit-exact on the whole frame; the tcgen05 CNN is
    checked against the oracle on four 256x320 crops (all four corners = all border cases) plus the centre."""
    import torch
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(2)
    img = _synth(rng, 1080, 1920)
    d_src = torch.from_numpy(img).to("cuda:0")
    pitch = 3840
    y, cr, cb = [torch.zeros((2160, pitch), dtype=torch.uint8, device="cuda:0") for _ in range(3)]
    engine.stage_color_bicubic(d_src, 2.0, y, cr, cb)
    yo = torch.zeros_like(y)
    engine.stage_cnn(y, yo, variant=S.VARIANT_TC)
    out = torch.zeros((2160, 3840, 3), dtype=torch.uint8, device="cuda:0")
    engine.stage_merge(yo, cr, cb, out)
    whole = torch.zeros_like(out)
    engine.process_device(d_src, 2.0, whole)
    engine.sync()
    assert torch.equal(out, whole)                       # staged == fused call
    ycc = oracle.bgr2ycrcb(img)
    up = [oracle.resize_cubic(ycc[:, :, k], 3840, 2160) for k in range(3)]
    assert np.array_equal(y.cpu().numpy(), up[0]) and np.array_equal(cr.cpu().numpy(), up[1]) and np.array_equal(cb.cpu().numpy(), up[2])
    Y = up[0]
    got = yo.cpu().numpy()
    H, W, ch, cw, m = 2160, 3840, 256, 320, 6
    for (r0, c0) in [(0, 0), (0, W - cw), (H - ch, 0), (H - ch, W - cw), (900, 1700)]:
        # crop of the upscaled Y with 6 px of context where the image continues; compare the part that does not
        # depend on what lies beyond the crop
        ra, rb, ca, cb_ = max(r0 - m, 0), min(r0 + ch + m, H), max(c0 - m, 0), min(c0 + cw + m, W)
        ref = oracle.cnn(np.ascontiguousarray(Y[ra:rb, ca:cb_]))
        ref = ref[r0 - ra:r0 - ra + ch, c0 - ca:c0 - ca + cw]
        st = diff_stats(got[r0:r0 + ch, c0:c0 + cw], ref)
        assert st["max"] <= 2 and st["le1"] >= 0.999, ((r0, c0), st)
    # merge stage bit-exact on the whole frame given the GPU's Y'
    assert np.array_equal(out.cpu().numpy(), oracle.ycrcb2bgr(np.dstack([got, up[1], up[2]])))


def test_cfg3_frame_batch_equals_single_frames(engine):
    """configs[2]: 1280x720 frames x2 -- a batch call equals frame-by-frame calls bit for bit (32 frames here)."""
    import torch
    rng = np.random.default_rng(3)
    base = _synth(rng, 720, 1280)
    frames = np.stack([np.roll(base, 7 * k, axis=1) for k in range(32)])
    d = torch.from_numpy(frames).to("cuda:0")
    batch = torch.zeros((32, 1440, 2560, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_batch_device(d, 2.0, batch)
    one = torch.zeros((1440, 2560, 3), dtype=torch.uint8, device="cuda:0")
    engine.sync()
    for k in (0, 13, 31):
        engine.process_device(d[k], 2.0, one)
        engine.sync()
        assert torch.equal(batch[k], one)
    # frame k is frame 0 rolled by 7k source pixels = 14k output pixels (away from the left/right borders)
    a, b = batch[0].cpu().numpy(), batch[5].cpu().numpy()
    assert np.array_equal(np.roll(a, 70, axis=1)[:, 200:-200], b[:, 200:-200])


def test_cfg4_row_bands_of_a_large_image(engine):
    """configs[3] (scaled to fit the test budget: 4096x2048 -> 8192x4096, 8 bands with the 6-px halo):
    band-by-band == whole, bit for bit, with each band seeing only its planned source rows."""
    import torch
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(4)
    img = _synth(rng, 2048, 4096)
    d_src = torch.from_numpy(img).to("cuda:0")
    whole = torch.zeros((4096, 8192, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_device(d_src, 2.0, whole)
    banded = torch.zeros_like(whole)
    edges = [4096 * i // 8 for i in range(9)]
    for r0, r1 in zip(edges[:-1], edges[1:]):
        s0, s1 = S.band_src_rows(2048, 2.0, r0, r1)
        assert s1 - s0 <= (r1 - r0) // 2 + 12          # <= 5 extra source rows per side at x2
        engine.process_band_device(d_src[s0:s1].contiguous(), 4096, 2048, s0, s1, 2.0, r0, r1, banded[r0:r1])
    engine.sync()
    assert torch.equal(whole, banded)


def test_cfg5_x4_frame_invariants(engine, oracle):
    """configs[4]: 3840x2160 -> 15360x8640 x4.  One full-size frame: output geometry, a flat image stays flat
    (interior of the CNN of a constant plane is a constant), and a crop matches the oracle."""
    import torch
    import srcnn_cpp_b200 as S
    assert S.out_dims(3840, 2160, 4.0) == (15360, 8640)
    flat = np.full((2160, 3840, 3), 90, np.uint8)
    d = torch.from_numpy(flat).to("cuda:0")
    out = torch.zeros((8640, 15360, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_device(d, 4.0, out)
    engine.sync()
    inner = out[16:-16, 16:-16]
    v = inner[0, 0].clone()
    assert bool((inner == v).all())
    # oracle on a small x4 case for the same scale tables (taps {-135,873,1535,-225}, ...)
    rng = np.random.default_rng(5)
    small = _synth(rng, 60, 80)
    got = engine.process(small, 4.0)
    st = diff_stats(got, oracle.pipeline(small, 4.0))
    assert st["max"] <= 2 and st["le1"] >= 0.999, st


def test_cfg4_full_size_32768_squared(engine):
    """configs[3] at its FULL size: one 32768x32768 -> 65536x65536 image (2^32 output pixels, 12.9 GB of BGR, 529 strips) on
    one GPU.  Checked through size-independent properties where 32-bit index arithmetic would break first:
      * a full-width row band far down the image (output rows 60000..60512, byte offsets beyond 2^33) == the same rows of the
        whole-image run, bit for bit;
      * a 400x400 source window deep inside the image, processed as an image of its own, == the whole-image result on the
        window's interior (the window starts on a source row that is a multiple of 11, so every output row keeps its ring
        slot and therefore its summation order in the tcgen05 kernel);
      * the four image corners are finite, non-constant data (no tile was skipped)."""
    import torch
    import srcnn_cpp_b200 as S
    free, _ = torch.cuda.mem_get_info()
    if free < 48 * 2**30:
        pytest.skip("needs ~36 GB of device memory")
    SW = SH = 32768
    OW = OH = 65536
    g = torch.Generator(device="cuda:0")
    g.manual_seed(44)
    # smooth-ish content: low-resolution noise repeated 64x in both directions plus fine grain (all on the device)
    base = torch.randint(0, 256, (SH // 64, SW // 64, 3), dtype=torch.uint8, device="cuda:0", generator=g)
    src = base.repeat_interleave(64, 0).repeat_interleave(64, 1)
    src += torch.randint(0, 24, (SH, SW, 3), dtype=torch.uint8, device="cuda:0", generator=g)   # wraps at 255: fine, still bytes
    del base
    whole = torch.empty((OH, OW, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_device(src, 2.0, whole)
    engine.sync()

    # (1) a band far down the image
    r0, r1 = 60000, 60512
    s0, s1 = S.band_src_rows(SH, 2.0, r0, r1)
    band = torch.zeros((r1 - r0, OW, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_band_device(src[s0:s1], SW, SH, s0, s1, 2.0, r0, r1, band)
    engine.sync()
    assert torch.equal(whole[r0:r1], band)

    # (2) a window deep inside, as an image of its own
    sy0, sx0, n = 11 * 2727, 31000, 400
    win = src[sy0:sy0 + n, sx0:sx0 + n].contiguous()
    out = torch.zeros((2 * n, 2 * n, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_device(win, 2.0, out)
    engine.sync()
    m = 16   # bicubic reach (4 output px) + the CNN's 6-px halo, rounded up
    ref = whole[2 * sy0 + m:2 * (sy0 + n) - m, 2 * sx0 + m:2 * (sx0 + n) - m]
    assert torch.equal(out[m:-m, m:-m], ref)

    # (3) corners
    for ys in (slice(0, 64), slice(OH - 64, OH)):
        for xs in (slice(0, 64), slice(OW - 64, OW)):
            c = whole[ys, xs].float()
            assert torch.isfinite(c).all() and float(c.std()) > 1.0
    del whole, src, band
    torch.cuda.empty_cache()


def test_cfg3_full_size_batch_of_1024_frames(engine):
    """configs[2] at its FULL size: 1024 frames of 1280x720 -> 2560x1440 in ONE batch call (2.8 GB in, 11.3 GB out).  Frame k is
    frame 0 rolled by 3 (k mod 97) source columns, so every result is known from frame 0's: the batch call must equal
    single-frame calls bit for bit on sampled frames (first, last, some in between), and the roll relation must hold away from
    the left/right borders for all sampled frames (a frame that was skipped, repeated or written to the wrong slot breaks it)."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 24 * 2**30:
        pytest.skip("needs ~16 GB of device memory")
    rng = np.random.default_rng(33)
    base = torch.from_numpy(_synth(rng, 720, 1280)).to("cuda:0")
    N = 1024
    shifts = [3 * (k % 97) for k in range(N)]
    src = torch.empty((N, 720, 1280, 3), dtype=torch.uint8, device="cuda:0")
    for k in range(N):
        src[k] = torch.roll(base, shifts[k], dims=1)
    dst = torch.zeros((N, 1440, 2560, 3), dtype=torch.uint8, device="cuda:0")
    engine.process_batch_device(src, 2.0, dst)
    engine.sync()
    one = torch.zeros((1440, 2560, 3), dtype=torch.uint8, device="cuda:0")
    for k in (0, 1, 511, 777, 1023):
        engine.process_device(src[k], 2.0, one)
        engine.sync()
        assert torch.equal(dst[k], one), k
    ref = dst[0]
    for k in range(0, N, 37):
        want = torch.roll(ref, 2 * shifts[k], dims=1)
        assert torch.equal(dst[k][:, 640:-640], want[:, 640:-640]), k
    del src, dst
    torch.cuda.empty_cache()
