"""CPU: the oracle's stand-alone metric-scan entry points (schro_metric_scan_setup / _do_scan / _get_min,
schro_metric_fast_block -- schroedinger/schrometric.c:31-214, 332-414) pinned bit-exactly against the
compiled, unmodified reference, and against golden vectors made from it."""
import ctypes
import os

import numpy as np
import pytest

from tests import helpers

ORACLE = helpers.load_oracle()
REF = helpers.load_ref()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (reference tree absent)")
GOLD = os.path.join(helpers.GOLDEN_DIR, "metric_scan.npz")
W, H = 96, 64

# (x, y, block_width, block_height, dx, dy, dist, use_chroma): interior, every picture edge, partial blocks,
# the largest window (41 x 41) and windows clipped by the 32-pixel border
QUERIES = [
    (32, 24, 8, 8, 0, 0, 3, 0), (32, 24, 8, 8, 2, -1, 3, 1), (0, 0, 8, 8, 0, 0, 3, 0), (0, 0, 8, 8, -5, -6, 5, 1),
    (88, 56, 8, 8, 3, 2, 3, 0), (88, 56, 8, 8, 9, 9, 10, 1), (40, 32, 8, 8, 0, 0, 20, 0), (40, 32, 8, 8, -4, 7, 20, 1),
    (92, 60, 4, 4, 1, 1, 3, 0), (48, 16, 12, 12, -3, 4, 5, 0), (16, 40, 16, 8, 6, -2, 4, 1), (64, 8, 8, 16, 0, 0, 7, 1),
    (8, 48, 8, 8, -14, 10, 10, 0), (80, 8, 8, 8, 14, -14, 20, 1),
]


def pictures(seed=3):
    return helpers.panning_pair(W, H, np.random.default_rng(seed), (3, -2), noise=6)


def levels(src, ref):
    ps = helpers.build_pyramid(ORACLE, "oracle", src, 0, ext_level0=32)
    pr = helpers.build_pyramid(ORACLE, "oracle", ref, 0, ext_level0=32)
    return helpers.pyr_level_struct(ps[0]), helpers.pyr_level_struct(pr[0]), (ps, pr)


def oracle_scan(src, ref, q):
    """-> (ref_x, ref_y, scan_w, scan_h, dx, dy, metric, chroma, fast_block), metrics, chroma_metrics"""
    x, y, bw, bh, dx, dy, dist, uc = q
    ls, lr, keep = levels(src, ref)
    rx, ry, sw, sh = (ctypes.c_int() for _ in range(4))
    ORACLE.oracle_metric_scan_setup(ctypes.byref(ls), x, y, bw, bh, dx, dy, dist, ctypes.byref(rx), ctypes.byref(ry),
                                    ctypes.byref(sw), ctypes.byref(sh))
    m = np.zeros(42 * 42, np.uint32)
    c = np.zeros(42 * 42, np.uint32)
    ORACLE.oracle_metric_scan_do_scan(ctypes.byref(ls), ctypes.byref(lr), x, y, bw, bh, rx.value, ry.value, sw.value,
                                      sh.value, uc, m.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p))
    odx, ody, chroma = ctypes.c_int(dx), ctypes.c_int(dy), ctypes.c_uint32()
    fn = ORACLE.oracle_metric_scan_get_min
    fn.restype = ctypes.c_uint32
    best = fn(m.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p), x, y, rx.value, ry.value, sw.value,
              sh.value, dx, dy, uc, ctypes.byref(odx), ctypes.byref(ody), ctypes.byref(chroma))
    fast = ORACLE.oracle_metric_fast_block(ctypes.byref(ls), ctypes.byref(lr), bw, bh, x, y, dx, dy)
    out = np.array([rx.value, ry.value, sw.value, sh.value, odx.value, ody.value, best, chroma.value, fast], np.int64)
    return out, m, c


def ref_scan(src, ref, q):
    P, I = ctypes.c_void_p * 3, ctypes.c_int * 3
    out = np.zeros(9, np.int32)
    m = np.zeros(42 * 42, np.uint32)
    c = np.zeros(42 * 42, np.uint32)
    qa = np.array(q, np.int32)
    fn = REF.ref_metric_scan
    fn.restype = None
    fn(W, H, P(*[a.ctypes.data for a in src]), I(*[a.strides[0] for a in src]), P(*[a.ctypes.data for a in ref]),
       I(*[a.strides[0] for a in ref]), qa.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
       m.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p))
    return out.astype(np.int64), m, c


def used(out):
    return int(out[2]) * int(out[3])


@needs_ref
def test_metric_scan_matches_reference():
    src, ref = pictures()
    for q in QUERIES:
        want, wm, wc = ref_scan(src, ref, q)
        got, gm, gc = oracle_scan(src, ref, q)
        assert np.array_equal(got, want), (q, got, want)
        n = used(want)
        assert np.array_equal(gm[:n], wm[:n]) and np.array_equal(gc[:n], wc[:n]), q
        assert wm[:n].min() == want[6] or q[7]


def test_metric_scan_golden():
    g = np.load(GOLD)
    src = [g[f"src{k}"] for k in range(3)]
    ref = [g[f"ref{k}"] for k in range(3)]
    assert len(g["queries"]) >= 12
    for i, q in enumerate(g["queries"]):
        got, gm, gc = oracle_scan(src, ref, tuple(int(v) for v in q))
        assert np.array_equal(got, g[f"q{i}_out"]), (i, got, g[f"q{i}_out"])
        n = used(got)
        assert np.array_equal(gm[:n], g[f"q{i}_metrics"]) and np.array_equal(gc[:n], g[f"q{i}_chroma"]), i
