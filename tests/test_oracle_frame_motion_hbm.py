"""CPU: the upsample / downsample / OBMC / block-matching oracles against the committed golden
vectors (outputs of the unmodified reference) and, where oracle/_ref exists, against the
reference itself on further cases."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.golden import make_golden as mg

ORACLE = helpers.load_oracle()
REF = helpers.load_ref()
G = {n: np.load(os.path.join(helpers.GOLDEN_DIR, n + ".npz")) for n in ("frame", "motion", "hbm")}


def test_upsample_and_edgeextend_golden():
    g = G["frame"]
    idx = 0
    while f"up{idx}_img" in g.files:
        img, ext = g[f"up{idx}_img"], int(g[f"up{idx}_ext"][0])
        pl = helpers.HostPlane(img.shape[1], img.shape[0], ext=ext, upsampled=True, fill=0x55)
        pl.set_image(img)
        helpers.cpu_edgeextend(ORACLE, "oracle", pl)
        helpers.cpu_upsample(ORACLE, "oracle", pl)
        for p in range(4):
            assert np.array_equal(pl.phase(p), g[f"up{idx}_phase{p}"]), (idx, p)
        idx += 1
    assert idx >= 6


def test_downsample_golden():
    g = G["frame"]
    idx = 0
    while f"down{idx}_img" in g.files:
        assert np.array_equal(helpers.cpu_downsample(ORACLE, "oracle", g[f"down{idx}_img"]),
                              g[f"down{idx}_out"]), idx
        idx += 1
    assert idx >= 6


def test_obmc_golden():
    g = G["motion"]
    for idx, kw in enumerate(mg.OBMC_GOLDEN_CASES):
        for add in (1, 0):
            if not add and kw.get("res_is_s32"):
                continue
            case = helpers.ObmcCase(ORACLE, rng=np.random.default_rng(1000 + idx), **kw)
            res = helpers.oracle_obmc(ORACLE, case, add)
            for k in range(3):
                for q, name in enumerate(("acc", "resid", "out")):
                    if q == 2 and not add:
                        continue
                    assert np.array_equal(res[k][q], g[f"c{idx}_add{add}_k{k}_{name}"]), (kw, add, k, name)


def test_hbm_golden():
    g = G["hbm"]
    for idx, (w, h, lv, uc, ri, pan) in enumerate(mg.HBM_GOLDEN_CASES):
        s, r = helpers.panning_pair(w, h, np.random.default_rng(2000 + idx), pan)
        fields, ps, _ = helpers.oracle_hbm(ORACLE, s, r, w, h, levels=lv, use_chroma=uc, ref_index=ri)
        want = g[f"h{idx}_fields"]
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(fields[f], want[f]), (idx, f)
        for l in range(lv):
            assert np.array_equal(ps[l + 1][0].phase(0, False), g[f"h{idx}_pyr{l}"]), (idx, l)
        # the synthetic pan is recovered on most blocks (sanity of the fixture itself)
        nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
        v = fields[0]["v"].reshape(nby, nbx, 4)
        inner = v[1:(h // 8) - 1, 1:(w // 8) - 1]
        hit = np.mean((inner[:, :, ri] == -pan[0]) & (inner[:, :, 2 + ri] == -pan[1]))
        if idx < 2:
            assert hit > 0.9, (idx, hit)


def test_sad_primitive():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, size=(40, 64)).astype(np.uint8)
    b = rng.integers(0, 256, size=(40, 64)).astype(np.uint8)
    import ctypes
    for (w, h) in ((8, 8), (12, 12), (16, 7), (32, 9), (5, 3), (1, 1)):
        want = int(np.abs(a[:h, :w].astype(int) - b[:h, :w].astype(int)).sum())
        ORACLE.oracle_sad_u8.restype = ctypes.c_uint32
        got = ORACLE.oracle_sad_u8(a.ctypes.data_as(ctypes.c_void_p), a.strides[0],
                                   b.ctypes.data_as(ctypes.c_void_p), b.strides[0], w, h)
        assert got == want
        if REF is not None:
            REF.ref_sad_u8.restype = ctypes.c_uint32
            assert REF.ref_sad_u8(a.ctypes.data_as(ctypes.c_void_p), a.strides[0],
                                  b.ctypes.data_as(ctypes.c_void_p), b.strides[0], w, h) == want


@pytest.mark.skipif(REF is None, reason="oracle/_ref not built (needs /root/reference)")
def test_frame_ops_match_reference_more_sizes():
    rng = np.random.default_rng(9)
    for (w, h, ext) in ((33, 12, 0), (100, 37, 32), (3, 9, 4), (2, 5, 1), (17, 17, 8), (8, 30, 5), (9, 9, 9)):
        img = rng.integers(0, 256, size=(h, w)).astype(np.uint8)
        outs = []
        for lib, pre in ((REF, "ref"), (ORACLE, "oracle")):
            pl = helpers.HostPlane(w, h, ext=ext, upsampled=True, fill=0xAA)
            pl.set_image(img)
            helpers.cpu_edgeextend(lib, pre, pl)
            helpers.cpu_upsample(lib, pre, pl)
            outs.append(pl.buf.copy())
        assert np.array_equal(outs[0], outs[1]), (w, h, ext)
    for w in range(1, 24):
        for h in (1, 2, 7, 16):
            img = rng.integers(0, 256, size=(h, w)).astype(np.uint8)
            assert np.array_equal(helpers.cpu_downsample(REF, "ref", img),
                                  helpers.cpu_downsample(ORACLE, "oracle", img)), (w, h)


@pytest.mark.skipif(REF is None, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kw", [
    dict(width=100, height=60, xbsep=8, ybsep=8, xblen=8, yblen=8),
    dict(width=64, height=48, weights=(1, 2, 2)),
    dict(width=64, height=48, num_refs=1, weights=(2, 1, 1)),
    dict(width=80, height=48, chroma_format=1),
    dict(width=352, height=288, span=200, outliers=0.05),
    dict(width=64, height=48, xbsep=8, ybsep=4, xblen=12, yblen=8, prec=3),
])
def test_obmc_matches_reference(kw):
    for add in (1, 0):
        case = helpers.ObmcCase(ORACLE, rng=np.random.default_rng(77), **kw)
        o = helpers.oracle_obmc(ORACLE, case, add)
        r = helpers.ref_obmc(REF, case, add)
        for k in range(3):
            for q in range(3 if add else 2):
                assert np.array_equal(o[k][q], r[k][q]), (kw, add, k, q)


@pytest.mark.skipif(REF is None, reason="oracle/_ref not built (needs /root/reference)")
def test_hbm_matches_reference_more_cases():
    for (w, h, lv, uc, ri, pan) in ((90, 50, 1, 0, 0, (1, 1)), (640, 360, 4, 0, 0, (5, 3)),
                                    (176, 144, 4, 1, 0, (12, 0)), (200, 120, 3, 0, 1, (0, -9))):
        s, r = helpers.panning_pair(w, h, np.random.default_rng(w), pan)
        fo, _, _ = helpers.oracle_hbm(ORACLE, s, r, w, h, levels=lv, use_chroma=uc, ref_index=ri)
        fr, _ = helpers.ref_hbm(REF, s, r, w, h, levels=lv, use_chroma=uc, ref_index=ri)
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(fo[f], fr[f]), (w, h, f)
