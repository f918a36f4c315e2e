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
