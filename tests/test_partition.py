"""CPU test of the row-walking kernel's work cut (srcnn_cpp_b200/csrc/srcnn_tc2.cu, tc2_partition): pure host
arithmetic reached through a debug export of the C-ABI library, no GPU needed.  The kernel's results do not depend on
the cut (tests/test_stage_parity.py::test_tc_result_independent_of_work_cut runs that on the GPU); here the cut itself
is checked: every row step of every strip is owned by exactly one pipeline, in order, and the most expensive pipeline is
never worse than with the equal-row-count cut it replaces."""
import ctypes as C

import numpy as np
import pytest

import srcnn_cpp_b200 as S


def _partition(nstrips, hb, nworkers, ovh):
    L = S.load_library()
    f = L.srcnn_debug_tc2_partition
    f.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_longlong)]
    f.restype = C.c_int
    b = (C.c_longlong * (nworkers + 1))()
    assert f(nstrips, hb, nworkers, ovh, b) == 0
    return np.array(b[:], dtype=np.int64)


def _cost(bounds, hb, ovh):
    """(cost of the most expensive pipeline, number of segments): a segment costs its rows + ovh row steps"""
    worst, nseg = 0, 0
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        c = 0
        while lo < hi:
            take = min(hi - lo, hb - lo % hb)
            c += take + ovh
            lo += take
            nseg += 1
        worst = max(worst, c)
    return worst, nseg


GEOMETRIES = [(31, 2160, 296), (21, 1440, 296), (31, 270, 296), (124, 8640, 296), (529, 8192, 296),
              (1, 50, 2), (3, 7, 4), (2, 96, 4), (1, 1, 2), (5, 3, 296), (31, 2160, 2), (7, 1000, 37)]


@pytest.mark.parametrize("nstrips,hb,nworkers", GEOMETRIES)
@pytest.mark.parametrize("ovh", [0, 4, 12, 16, 64])
def test_cut_covers_every_row_step_once(nstrips, hb, nworkers, ovh):
    b = _partition(nstrips, hb, nworkers, ovh)
    assert b[0] == 0 and b[-1] == nstrips * hb
    assert (np.diff(b) >= 0).all()


@pytest.mark.parametrize("nstrips,hb,nworkers", [g for g in GEOMETRIES if g[0] * g[1] >= 24 * g[2]])
def test_cost_aware_cut_is_never_worse_than_equal_rows(nstrips, hb, nworkers):
    # (launches give every pipeline ~48 row steps or use fewer CTAs: launch_cnn_tc2)
    ovh = 12
    even, _ = _cost(_partition(nstrips, hb, nworkers, 0), hb, ovh)
    aware, _ = _cost(_partition(nstrips, hb, nworkers, ovh), hb, ovh)
    assert aware <= even


def test_equal_rows_cut_is_the_plain_formula():
    b = _partition(31, 2160, 296, 0)
    assert all(int(b[w]) == 31 * 2160 * w // 296 for w in range(297))


def test_bench_geometry_gain():
    """1080p -> 4K: 31 strips x 2160 rows over 296 pipelines; the slowest pipeline drops from 251 to <= 241 row steps"""
    even, _ = _cost(_partition(31, 2160, 296, 0), 2160, 12)
    aware, _ = _cost(_partition(31, 2160, 296, 12), 2160, 12)
    assert even >= 250 and aware <= 241


def test_bad_arguments_are_refused():
    L = S.load_library()
    f = L.srcnn_debug_tc2_partition
    f.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_longlong)]
    b = (C.c_longlong * 4)()
    assert f(0, 10, 2, 12, b) != 0 and f(1, 0, 2, 12, b) != 0 and f(1, 10, 0, 12, b) != 0
