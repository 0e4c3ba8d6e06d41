"""GPU parity: sub-pel refinement of the motion fields (sb2_subpel_refine, the
schro_b200_motion_predict_subpel_deep drop-in and the reference-side shim of
schro_encoder_motion_predict_subpel_deep) against the oracle, bit-exact -- including the double-precision
score comparisons."""
import ctypes
import os

import numpy as np
import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()
SHIM = os.path.join(helpers.ROOT, "oracle", "_ref", "libcompat_shim.so")


@pytest.fixture(params=["words", "pixels"], autouse=True)
def probe_path(request):
    """Every test runs twice: full 8 x 8 blocks through the word-wide probe path (default) and through the
    per-pixel path every other block takes (sb2_subpel_force_generic)."""
    from schroedinger_b200 import lib
    lib.sb2_subpel_force_generic(1 if request.param == "pixels" else 0)
    yield request.param
    lib.sb2_subpel_force_generic(0)


def jitter(fields, rng, frac=3):
    for r, f in enumerate(fields):
        idx = rng.choice(len(f), len(f) // frac, replace=False)
        f["v"][idx, r] += rng.integers(-1, 2, size=len(idx)).astype(np.int16)
        f["v"][idx, 2 + r] += rng.integers(-1, 2, size=len(idx)).astype(np.int16)


def gpu_subpel(cases, w, h, bs, prec, lam):
    """cases: list of (src, refs, fields); pictures run as one batch per reference index."""
    from schroedinger_b200 import device as dev
    count = len(cases)
    nbx, nby = helpers.hbm_block_counts(w, h, bs, bs)
    orig = dev.PictureSlab(dev.FrameLayout.yuv420("u8", w, h, 32), count)
    nrefs = len(cases[0][1])
    out = [[None] * nrefs for _ in cases]
    for p, (src, refs, fields) in enumerate(cases):
        for c in range(3):
            orig.upload(p, c, src[c])
    for r in range(nrefs):
        up = dev.PictureSlab(dev.FrameLayout.yuv420("u8", w, h, 32, True), count)
        for p, (src, refs, fields) in enumerate(cases):
            for c in range(3):
                up.upload(p, c, refs[r][c])
        dev.edgeextend_upsample(up)
        fld = torch.from_numpy(np.concatenate([c[2][r] for c in cases]).view(np.uint8).copy()).cuda()
        dev.subpel_refine(orig, up, fld, bs, bs, nbx, nby, prec, r, lam)
        got = fld.cpu().numpy().view(helpers.MV_DTYPE).reshape(count, nbx * nby)
        for p in range(count):
            out[p][r] = got[p]
    return out


def check(got, want, what):
    for r in range(len(want)):
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(got[r][f], want[r][f]), (what, r, f)


@pytest.mark.parametrize("prec", [1, 2, 3])
@pytest.mark.parametrize("lam", [0.0, 0.1, 1.0, 10.0])
def test_subpel_vs_oracle(cuda, prec, lam):
    w, h = 176, 144
    rng = np.random.default_rng(prec * 10 + int(lam))
    src, refs, fields = helpers.subpel_case(ORACLE, w, h, rng)
    jitter(fields, rng)
    want = helpers.oracle_subpel(ORACLE, src, refs, fields, w, h, 8, 8, prec, lam)
    got = gpu_subpel([(src, refs, fields)], w, h, 8, prec, lam)
    check(got[0], want, (prec, lam))
    assert any(np.any(want[r]["v"] != (fields[r]["v"].astype(np.int32) << prec)) for r in range(2))


@pytest.mark.parametrize("w,h,bs,prec", [(200, 104, 8, 2), (100, 70, 8, 3), (192, 144, 12, 2), (256, 128, 16, 2), (64, 48, 4, 2)])
def test_subpel_ragged_sizes_and_block_sizes(cuda, w, h, bs, prec):
    rng = np.random.default_rng(w + h + bs)
    src, refs, _ = helpers.subpel_case(ORACLE, w, h, rng)
    nbx, nby = helpers.hbm_block_counts(w, h, bs, bs)
    fields = [np.zeros(nbx * nby, helpers.MV_DTYPE) for _ in refs]
    for r, f in enumerate(fields):
        f["flags"] = r + 1
        f["v"][:, r] = rng.integers(-6, 7, size=nbx * nby)
        f["v"][:, 2 + r] = rng.integers(-6, 7, size=nbx * nby)
        f["metric"] = rng.integers(0, 3000, size=nbx * nby)
    want = helpers.oracle_subpel(ORACLE, src, refs, fields, w, h, bs, bs, prec, 0.07)
    got = gpu_subpel([(src, refs, fields)], w, h, bs, prec, 0.07)
    check(got[0], want, (w, h, bs))


def test_subpel_batch(cuda):
    """Several pictures per launch: every picture's decision wavefront is its own CTA."""
    w, h = 320, 192
    rng = np.random.default_rng(99)
    cases = []
    for p in range(5):
        src, refs, fields = helpers.subpel_case(ORACLE, w, h, rng, pans=((p - 2, 3 - p), (2 * p - 3, -p)))
        jitter(fields, rng, frac=2)
        cases.append((src, refs, fields))
    got = gpu_subpel(cases, w, h, 8, 2, 0.1)
    for p, (src, refs, fields) in enumerate(cases):
        want = helpers.oracle_subpel(ORACLE, src, refs, fields, w, h, 8, 8, 2, 0.1)
        check(got[p], want, p)


def test_subpel_1080p(cuda):
    """BASELINE-sized picture: 240 x 136 blocks, quarter-pel, vs the oracle."""
    w, h = 1920, 1080
    rng = np.random.default_rng(7)
    src, rf = helpers.panning_pair(w, h, rng, (5, 3))
    nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
    f = np.zeros(nbx * nby, helpers.MV_DTYPE)
    f["flags"] = 1
    f["v"][:, 0] = -5 + rng.integers(-1, 2, size=nbx * nby)
    f["v"][:, 2] = -3 + rng.integers(-1, 2, size=nbx * nby)
    f["metric"] = rng.integers(100, 2000, size=nbx * nby)
    want = helpers.oracle_subpel(ORACLE, src, [rf], [f], w, h, 8, 8, 2, 0.1)
    got = gpu_subpel([(src, [rf], [f])], w, h, 8, 2, 0.1)
    check(got[0], want, "1080p")


def _frames_and_fields(compat, lib, w, h, src, refs, fields):
    from tests.test_host_api_gpu import _new_u8_frame
    fo = _new_u8_frame(compat, lib, w, h, 32, False, src)
    ups = []
    for rf in refs:
        u = _new_u8_frame(compat, lib, w, h, 32, True, rf)
        lib.schro_frame_mc_edgeextend(u)
        ups.append(u)
    mfs = []
    for f in fields:
        mf = lib.schro_motion_field_new(compat_nbx[0], compat_nbx[1])
        ctypes.memmove(mf.contents.motion_vectors, f.ctypes.data, f.nbytes)
        mfs.append(mf)
    return fo, ups, mfs


compat_nbx = [0, 0]


@pytest.mark.parametrize("via_shim", [False, True])
def test_subpel_drop_in(cuda, via_shim):
    from schroedinger_b200 import compat, lib
    if via_shim and not os.path.exists(SHIM):
        pytest.skip("oracle/_ref/libcompat_shim.so was not built")
    w, h, prec, lam = 176, 144, 2, 0.1
    rng = np.random.default_rng(41)
    src, refs, fields = helpers.subpel_case(ORACLE, w, h, rng)
    jitter(fields, rng)
    want = helpers.oracle_subpel(ORACLE, src, refs, fields, w, h, 8, 8, prec, lam)
    params = compat.make_params(w, h, xbsep=8, ybsep=8, xblen=12, yblen=12)
    params.num_refs = 2
    params.mv_precision = prec
    compat_nbx[0], compat_nbx[1] = params.x_num_blocks, params.y_num_blocks
    fo, ups, mfs = _frames_and_fields(compat, lib, w, h, src, refs, fields)
    FP = compat.FrameP * 2
    MP = ctypes.POINTER(compat.SchroMotionField) * 2
    if via_shim:
        from schroedinger_b200._lib import LIB_PATH
        ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        shim = ctypes.CDLL(SHIM)
        shim.compat_shim_subpel_deep.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        shim.compat_shim_subpel_deep(ctypes.byref(params), lam, fo, FP(*ups), MP(*mfs))
    else:
        lib.schro_b200_motion_predict_subpel_deep(ctypes.byref(params), lam, fo, FP(*ups), MP(*mfs))
    n = params.x_num_blocks * params.y_num_blocks
    for r in range(2):
        got = np.ctypeslib.as_array(ctypes.cast(mfs[r].contents.motion_vectors, ctypes.POINTER(ctypes.c_uint8)),
                                    shape=(n * 20,)).view(helpers.MV_DTYPE)
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(got[f], want[r][f]), (r, f)
        assert ups[r].contents.upsample_done == 1
        lib.schro_motion_field_free(mfs[r])
        lib.schro_frame_unref(ups[r])
    lib.schro_frame_unref(fo)
