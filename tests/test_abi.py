"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "srcnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srcnn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import srcnn_cpp_b200 as S
    L = S.load_library()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libsrcnn_b200.so does not export " + n
    assert sorted(S.ABI_SYMBOLS) == names
    assert L.srcnn_abi_version() == 2


def test_out_dims_truncation():
    import srcnn_cpp_b200 as S
    assert S.out_dims(384, 384, 1.5) == (576, 576)
    assert S.out_dims(1920, 1080, 2.0) == (3840, 2160)
    assert S.out_dims(37, 29, 1.5) == (55, 43)       # 55.5 -> 55, 43.5 -> 43 (truncation, src/srcnn.cpp:573-575)
    assert S.out_dims(3, 9, 2.7) == (8, 24)
    with pytest.raises(S.SrcnnError) as e:
        S.out_dims(4, 4, 0.1)
    assert e.value.status == S.E_RATIO


def test_band_src_rows_cover_the_taps(oracle):
    import srcnn_cpp_b200 as S
    for (h, scale) in [(1080, 2.0), (720, 2.0), (2160, 4.0), (384, 1.5), (100, 3.0)]:
        oh = int(np.float32(h) * np.float32(scale))
        ofs, _ = oracle.cubic_taps(h, oh)
        for (r0, r1) in [(0, oh), (0, oh // 3), (oh // 3, 2 * oh // 3), (oh - 7, oh), (5, 6)]:
            s0, s1 = S.band_src_rows(h, scale, r0, r1)
            p0, p1 = max(r0 - 6, 0), min(r1 + 6, oh)
            need0 = min(max(int(ofs[p0]) - 1, 0), h - 1)
            need1 = min(max(int(ofs[p1 - 1]) + 2, 0), h - 1) + 1
            assert (s0, s1) == (need0, need1)
            assert 0 <= s0 < s1 <= h


def test_no_gpu_means_loud_failure():
    """The product has no CPU path: without a usable device the context cannot be created."""
    import torch
    import srcnn_cpp_b200 as S
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(S.SrcnnError) as e:
        S.Engine(device=0)
    assert e.value.status == S.E_NODEVICE


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "srcnn_cpp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in txt.replace("oracle/:", "") or f == "__init__.py" and "import oracle" not in txt, f
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_fused_kernel_builds_without_local_memory():
    """A register spill around an in-flight tcgen05.ld would store a register the TMEM load has not written yet
    (DESIGN.md section 4): the production instance of the row-walking kernel must not use local memory at all, and it
    must be sm_100a tcgen05 code (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = TMA bulk copy)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "srcnn_cpp_b200", "libsrcnn_b200.so")
    res = subprocess.run([cuobjdump, "-res-usage", lib], capture_output=True, text=True).stdout
    for inst in ("ILb0ELb0", "ILb0ELb1"):    # <DBG=false, FUSED=false|true>: the two production instances
        m = re.search(r"Function \S*k_srcnn_tc2" + inst + r"\S*:\s*\n\s*REG:(\d+) STACK:(\d+)", res)
        assert m, "k_srcnn_tc2<%s> not found in the library" % inst
        assert int(m.group(2)) == 0, "k_srcnn_tc2<%s> uses local memory (stack %s bytes)" % (inst, m.group(2))
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN5srcnn3tc211k_srcnn_tc2ILb0ELb1EEEvNS0_6ParamsE", lib],
                          capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR"):
        assert mnemonic in sass, mnemonic + " missing from k_srcnn_tc2's SASS"
    assert "BRA.U.ANY" not in sass.split("UTCHMMA", 1)[1], "a tcgen05.mma sits in a waterfall (non-uniform operand) loop"
