"""GPU tests of the sharding contract: row bands with the 6-px halo and frame batches
are bit-identical to the unsplit result, for both variants."""
import numpy as np
import pytest

from conftest import natural_like

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant_name", ["fp32", "tc"])
@pytest.mark.parametrize("w,h,scale,nb", [(60, 90, 2.0, 4), (41, 77, 3.0, 3), (48, 70, 1.5, 5), (40, 200, 2.0, 3)])
def test_bands_equal_whole(engine, variant_name, w, h, scale, nb):
    import torch
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(w * h)
    img = natural_like(rng, h, w)
    ow, oh = S.out_dims(w, h, scale)
    engine.set_variant(S.VARIANT_FP32 if variant_name == "fp32" else S.VARIANT_TC)
    try:
        d_src = torch.from_numpy(img).to("cuda:0")
        whole = torch.zeros((oh, ow, 3), dtype=torch.uint8, device="cuda:0")
        engine.process_device(d_src, scale, whole)
        engine.sync()
        banded = torch.zeros_like(whole)
        edges = [oh * i // nb for i in range(nb + 1)]
        for r0, r1 in zip(edges[:-1], edges[1:]):
            s0, s1 = S.band_src_rows(h, scale, r0, r1)
            src_rows = d_src[s0:s1].contiguous()      # this band's GPU only ever sees these rows
            engine.process_band_device(src_rows, w, h, s0, s1, scale, r0, r1, banded[r0:r1])
        engine.sync()
    finally:
        engine.set_variant(S.VARIANT_TC)
    assert torch.equal(whole, banded)


@pytest.mark.parametrize("w,h,scale,n", [(70, 90, 1.5, 5), (130, 64, 2.0, 3), (64, 48, 3.0, 4), (96, 80, 0.5, 3)])
def test_batch_in_one_launch_equals_frame_by_frame(engine, w, h, scale, n):
    """srcnn_process_batch_device runs a chunk of frames as ONE launch per stage (frames = a grid dimension of the colour+bicubic
    kernel, more strips of the row-walking kernel's work list, one tall image for the merge kernel): bit-identical to calling
    the frames one by one, also when the caller's result frames are not contiguous, and for a geometry the tiled kernel does not
    take (down-scale: direct kernel + separate FP16 pass)."""
    import ctypes as C
    import torch
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(w + h + n)
    frames = np.stack([natural_like(rng, h, w) for _ in range(n)])
    ow, oh = S.out_dims(w, h, scale)
    d = torch.from_numpy(frames).to("cuda:0")
    one = torch.zeros((n, oh, ow, 3), dtype=torch.uint8, device="cuda:0")
    for k in range(n):
        engine.process_device(d[k], scale, one[k])
    batch = torch.zeros_like(one)
    engine.process_batch_device(d, scale, batch)
    gappy = torch.zeros((n, oh + 3, ow, 3), dtype=torch.uint8, device="cuda:0")[:, :oh]     # frame stride != rows x row stride
    engine.process_batch_device(d, scale, gappy)
    engine.L.srcnn_debug_set_batch_launch.argtypes = [C.c_void_p, C.c_int]
    engine.L.srcnn_debug_set_batch_launch(engine.ctx, 0)
    try:
        loop = torch.zeros_like(one)
        engine.process_batch_device(d, scale, loop)
    finally:
        engine.L.srcnn_debug_set_batch_launch(engine.ctx, 1)
    engine.sync()
    assert torch.equal(batch, one) and torch.equal(gappy, one) and torch.equal(loop, one)


def test_batch_after_dirty_workspace(engine):
    """A batch whose work list crosses frame boundaries inside one pipeline, run on a workspace that earlier, larger calls left
    full of 0xFF bytes (NaN as FP16).  Lanes of a frame's last strip that lie beyond the image stage whatever the FP16 plane
    holds past its replicated columns; the next segment of the same pipeline (strip 0 of the next frame) must not inherit those
    bits through the im2col ring's zero-weight slots (0 x NaN = NaN in its first rows).  Regression test: found with frames of
    160x90 after 700x400 frames, three rows of two frames wrong."""
    import torch
    rng = np.random.default_rng(77)
    white = torch.from_numpy(rng.integers(250, 256, (400, 700, 3), dtype=np.uint8)).to("cuda:0")
    big_out = torch.zeros((800, 1400, 3), dtype=torch.uint8, device="cuda:0")
    import srcnn_cpp_b200 as S
    # (w, h, scale, frames): the geometry of the find; one strip exactly (2 lanes beyond the image); a last strip of 4 columns;
    # a x3 up-scale and a width that is no multiple of 8 (generic colour+bicubic kernel)
    for w, h, scale, n in ((160, 90, 2.0, 6), (160, 90, 2.0, 4), (160, 90, 2.0, 5), (160, 90, 2.0, 3), (62, 40, 2.0, 9), (250, 70, 2.0, 5),
                           (100, 60, 3.0, 4), (333, 50, 2.0, 3)):
        ow, oh = S.out_dims(w, h, scale)
        for _ in range(2):                                   # both Cr/Cb pairs
            engine.process_device(white, 2.0, big_out)
        frames = torch.from_numpy(np.stack([natural_like(rng, h, w) for _ in range(n)])).to("cuda:0")
        one = torch.zeros((n, oh, ow, 3), dtype=torch.uint8, device="cuda:0")
        for k in range(n):
            engine.process_device(frames[k], scale, one[k])
        for _ in range(2):
            engine.process_device(white, 2.0, big_out)
        for rep in range(2):
            batch = torch.zeros_like(one)
            engine.process_batch_device(frames, scale, batch)
            engine.sync()
            assert torch.equal(batch, one), (w, h, scale, n, rep)
