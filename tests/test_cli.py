"""bin/srcnn drop-in: argument handling and exit codes on CPU; the cfg1 golden run on the GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin", "srcnn")


def _run(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True, timeout=300)


def test_help_and_no_args_return_zero():
    for args in ((), ("--help",), ("--help", "x.png")):
        r = _run(*args)
        assert r.returncode == 0                                   # src/srcnn.cpp:709-715
        assert "usage :" in r.stdout and "--scale=" in r.stdout and "--noverbose" in r.stdout


def test_missing_source_is_exit_minus_one(tmp_path):
    r = _run(str(tmp_path / "nope.png"))
    assert r.returncode == 255                                     # t_exit_code = -1, src/srcnn.cpp:479
    assert "- load failure :" in r.stdout
    r = _run("--noverbose", str(tmp_path / "nope.png"))
    assert r.returncode == 255 and r.stdout == ""


@pytest.mark.gpu
def test_cfg1_butterfly_via_cli_is_bit_exact(tmp_path):
    """BASELINE configs[0]: Pictures/butterfly.png x1.5 via bin/srcnn == Pictures/butterfly-srcnn.png."""
    import cv2
    src = os.path.join(ROOT, "tests", "golden", "butterfly.png")
    gold = cv2.imread(os.path.join(ROOT, "tests", "golden", "butterfly-srcnn.png"))
    dst = str(tmp_path / "out.png")
    r = _run("--scale=1.5", "--variant=fp32", src, dst)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "- Performace :" in r.stdout and "ms took." in r.stdout
    assert np.array_equal(cv2.imread(dst), gold)
    # tensor-core variant + default output name <stem>_resized<ext> (src/srcnn.cpp:396-416)
    src2 = str(tmp_path / "b.png")
    cv2.imwrite(src2, cv2.imread(src))
    r = _run("--scale=1.5", "--noverbose", src2)
    assert r.returncode == 0 and r.stdout == ""
    out = cv2.imread(str(tmp_path / "b_resized.png"))
    d = np.abs(out.astype(np.int16) - gold.astype(np.int16))
    assert d.max() <= 2 and (d <= 1).mean() >= 0.999


@pytest.mark.gpu
def test_process_srcnn_library_entry(oracle):
    """ProcessSRCNN(rgb, w, h, d, mul, out&, outsz&) -- src/test.cpp:347-361: ret == 0 and
    outsz == (unsigned)(w*mul) * (unsigned)(h*mul) * d; RGB order in and out."""
    import srcnn_cpp_b200 as S
    L = S.load_library()
    fn = L._Z12ProcessSRCNNPKhjjjfRPhRj
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_uint)]
    rng = np.random.default_rng(4)
    bgr = rng.integers(0, 256, (20, 26, 3), dtype=np.uint8)
    rgb = np.ascontiguousarray(bgr[:, :, ::-1])
    out, sz = C.c_void_p(), C.c_uint()
    assert fn(rgb.ctypes.data, 26, 20, 3, C.c_float(2.0), C.byref(out), C.byref(sz)) == 0
    assert sz.value == 52 * 40 * 3
    got = np.ctypeslib.as_array((C.c_uint8 * sz.value).from_address(out.value)).reshape(40, 52, 3)
    want = oracle.pipeline(bgr, 2.0)[:, :, ::-1]
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= 2 and (d <= 1).mean() >= 0.999
    # 4-channel input keeps 4 channels (src/test.cpp: outsz uses d)
    rgba = np.dstack([rgb, np.full((20, 26), 200, np.uint8)])
    assert fn(np.ascontiguousarray(rgba).ctypes.data, 26, 20, 4, C.c_float(2.0), C.byref(out), C.byref(sz)) == 0
    assert sz.value == 52 * 40 * 4


@pytest.mark.gpu
def test_jpeg_in_and_out(tmp_path, oracle):
    """README.md:96-100 usage: ./bin/srcnn ./Pictures/test.jpg -> test_resized.jpg (JPEG through nvJPEG; file
    codecs are outside the parity contract, so this checks geometry and a PSNR against the oracle on the
    cv2-decoded input)."""
    import cv2
    img = cv2.imread(os.path.join(ROOT, "tests", "golden", "butterfly.png"))[:96, :128]
    src = str(tmp_path / "t.jpg")
    cv2.imwrite(src, img, [cv2.IMWRITE_JPEG_QUALITY, 95])
    r = _run("--noverbose", src)
    assert r.returncode == 0, r.stdout + r.stderr
    out = cv2.imread(str(tmp_path / "t_resized.jpg"))
    assert out is not None and out.shape == (192, 256, 3)
    want = oracle.pipeline(cv2.imread(src), 2.0)
    mse = np.mean((out.astype(np.float64) - want.astype(np.float64)) ** 2)
    assert 10 * np.log10(255.0 ** 2 / mse) > 30.0


@pytest.mark.gpu
def test_devices_option_splits_the_image_into_bands(tmp_path):
    """--devices=a,b,.. : one row band per listed device through srcnn_mgpu_process_banded_host; same bytes as one device
    (a device may be listed twice, so this runs on a one-GPU box too)."""
    import cv2
    src = os.path.join(ROOT, "tests", "golden", "butterfly.png")
    gold = cv2.imread(os.path.join(ROOT, "tests", "golden", "butterfly-srcnn.png"))
    one, many = str(tmp_path / "one.png"), str(tmp_path / "many.png")
    assert _run("--scale=1.5", "--variant=fp32", "--noverbose", src, one).returncode == 0
    r = _run("--scale=1.5", "--variant=fp32", "--devices=0,0,0", src, many)
    assert r.returncode == 0, r.stdout + r.stderr
    assert np.array_equal(cv2.imread(many), cv2.imread(one)) and np.array_equal(cv2.imread(many), gold)
    r = _run("--scale=1.5", "--devices=all", "--noverbose", src, many)
    assert r.returncode == 0
    d = np.abs(cv2.imread(many).astype(np.int16) - gold.astype(np.int16))
    assert d.max() <= 2 and (d <= 1).mean() >= 0.999
    r = _run("--devices=0,99", "--noverbose", src, many)      # a device that does not exist
    assert r.returncode != 0


def _png(w, h, depth, ctype, rows, palette=None):
    """Hand-made PNG (filter 0 on every row): rows = list of packed scanlines."""
    import struct
    import zlib

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    raw = b"".join(b"\x00" + bytes(r) for r in rows)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if palette is not None:
        out += chunk(b"PLTE", bytes(palette))
    return out + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")


def _sub_byte_cases():
    rng = np.random.default_rng(21)
    w, h = 37, 23                                     # not a multiple of 8, 4 or 2 pixels per byte
    cases = []
    for depth in (1, 2, 4):
        vals = rng.integers(0, 1 << depth, (h, w), dtype=np.uint8)
        per = 8 // depth
        rows = []
        for y in range(h):
            row = bytearray((w + per - 1) // per)
            for x in range(w):
                row[x // per] |= int(vals[y, x]) << ((per - 1 - x % per) * depth)
            rows.append(row)
        grey = (vals.astype(np.int32) * (255 // ((1 << depth) - 1))).astype(np.uint8)
        cases.append(("grey%d" % depth, _png(w, h, depth, 0, rows), np.dstack([grey] * 3)))
        pal = rng.integers(0, 256, (1 << depth, 3), dtype=np.uint8)          # R,G,B triples
        cases.append(("pal%d" % depth, _png(w, h, depth, 3, rows, pal.reshape(-1)), pal[vals][:, :, ::-1]))   # -> BGR
    return cases


def test_png_sub_byte_depths_and_bad_headers_on_cpu(tmp_path):
    """1/2/4-bit grey and palette PNGs decode (cv::imread accepts them, src/srcnn.cpp:462); a header with absurd dimensions
    or a truncated stream is a load failure (exit -1), never a crash.  Without a GPU the run ends at context creation --
    after the '- Image load' line, which is what this checks."""
    import struct
    for name, data, _ in _sub_byte_cases():
        f = tmp_path / (name + ".png")
        f.write_bytes(data)
        r = _run(str(f), str(tmp_path / "o.png"))
        assert "- Image load : " in r.stdout and "load failure" not in r.stdout, (name, r.stdout)
    good = _sub_byte_cases()[0][1]
    huge = bytearray(good)
    huge[16:24] = struct.pack(">II", 0x7FFFFFFF, 0x7FFFFFFF)       # IHDR width / height (CRC is not checked by the reader)
    for name, data in (("huge", bytes(huge)), ("cut", good[:len(good) - 30]), ("junk", b"\x89PNG\r\n\x1a\n" + b"\x00" * 40)):
        f = tmp_path / (name + ".png")
        f.write_bytes(data)
        r = _run(str(f), str(tmp_path / "o.png"))
        assert r.returncode == 255 and "- load failure :" in r.stdout, (name, r.returncode, r.stdout)


def test_scale_values_that_are_ignored(tmp_path):
    """--scale=0, a negative or an unparsable ratio leave the default 2.0 in place (src/srcnn.cpp:359-370)."""
    for arg in ("--scale=0", "--scale=-3", "--scale=abc", "--scale="):
        r = _run(arg, str(tmp_path / "nope.png"))
        assert "- Scale multiply ratio : 2.00" in r.stdout and r.returncode == 255
    r = _run("--scale=1.5", "--devices=0,1,x,,3", str(tmp_path / "nope.png"))
    assert "- Scale multiply ratio : 1.50" in r.stdout and r.returncode == 255


@pytest.mark.gpu
def test_png_sub_byte_depths_decode_like_opencv(tmp_path, oracle):
    """The decoded pixels are the ones cv::imread would hand the pipeline: FP32 variant == oracle on the expected BGR image."""
    import cv2
    for name, data, bgr in _sub_byte_cases():
        f = tmp_path / (name + ".png")
        f.write_bytes(data)
        assert np.array_equal(cv2.imread(str(f)), bgr), name                 # the test's expectation is OpenCV's decode
        out = str(tmp_path / (name + "_out.png"))
        r = _run("--variant=fp32", "--noverbose", str(f), out)
        assert r.returncode == 0, (name, r.stdout)
        assert np.array_equal(cv2.imread(out), oracle.pipeline(np.ascontiguousarray(bgr), 2.0)), name
