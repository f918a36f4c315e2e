"""Stream ingest: srcnn_jpeg_stream_* -- JPEG frames in, JPEG frames out, decode / kernels / encode overlapped on
three streams with a ring of three device frames.  Codecs are outside the parity contract (IDCT rounding differs between
decoders), so the check is geometry, frame order and a PSNR against the oracle run on the cv2-decoded input."""
import numpy as np
import pytest

from conftest import natural_like

pytestmark = pytest.mark.gpu


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 10 * np.log10(255.0 ** 2 / max(mse, 1e-9))


def test_stream_of_frames(engine, oracle):
    import cv2
    import srcnn_cpp_b200 as S
    rng = np.random.default_rng(8)
    n, h, w = 7, 96, 128                                     # more frames than ring slots
    frames = [natural_like(rng, h, w) for _ in range(n)]
    for k, f in enumerate(frames):
        f[:8, :8] = 30 * k                                   # a tag per frame: order must survive the pipeline
    jpegs = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes() for f in frames]
    st = S.JpegStream(engine, quality=95)
    try:
        outs, (ow, oh) = st.process(jpegs, 2.0)
        assert (ow, oh) == (2 * w, 2 * h) and len(outs) == n
        for k in range(n):
            got = cv2.imdecode(np.frombuffer(outs[k], np.uint8), cv2.IMREAD_COLOR)
            assert got is not None and got.shape == (oh, ow, 3)
            want = oracle.pipeline(cv2.imdecode(np.frombuffer(jpegs[k], np.uint8), cv2.IMREAD_COLOR), 2.0)
            assert _psnr(got, want) > 30.0, (k, _psnr(got, want))
            assert abs(float(got[:10, :10].mean()) - float(want[:10, :10].mean())) < 6.0   # this frame's tag, not a neighbour's
        # a second call reuses the ring; an empty call is fine
        outs2, _ = st.process(jpegs[:2], 2.0)
        assert outs2[0] == outs[0] and outs2[1] == outs[1]
        assert st.process([], 2.0)[0] == []
        # a frame of another size, or garbage, is an argument error -- and the stream stays usable
        other = cv2.imencode(".jpg", natural_like(rng, 64, 64))[1].tobytes()
        for bad in ([jpegs[0], other], [jpegs[0], b"not a jpeg at all"], [b"\xff\xd8\xff junk"]):
            with pytest.raises(S.SrcnnError) as e:
                st.process(bad, 2.0)
            assert e.value.status == S.E_ARG
        assert st.process(jpegs[:1], 2.0)[0][0] == outs[0]
    finally:
        st.close()
