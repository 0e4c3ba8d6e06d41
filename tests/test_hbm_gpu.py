"""GPU parity: pyramid + hierarchical block matching (sb2_downsample, sb2_mc_edgeextend,
sb2_hbm_scan_hint) and the SAD primitive against the oracle / golden fields, bit-exact."""
import os

import numpy as np
import pytest
import torch

from tests import helpers
from tests.golden import make_golden as mg

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()
GOLD = np.load(os.path.join(helpers.GOLDEN_DIR, "hbm.npz"))


@pytest.fixture(params=["wave", "generic"], autouse=True)
def hbm_kernel(request):
    """Every test of this file runs twice: with the kernel the library picks (the skewed wavefront
    for 8x8 / 4:2:0 / luma-only scans) and with the generic one-row-per-CTA kernel forced."""
    from schroedinger_b200 import lib
    lib.sb2_hbm_force_generic(1 if request.param == "generic" else 0)
    yield request.param
    lib.sb2_hbm_force_generic(0)


def launched_tags(fn):
    """Run fn with per-launch profiling on; return the set of kernel tags it launched."""
    import ctypes
    from schroedinger_b200 import lib
    lib.sb2_profile_reset()
    lib.sb2_profile_enable(1)
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        lib.sb2_profile_enable(0)
    tags = set()
    buf = ctypes.create_string_buffer(64)
    ms, by = ctypes.c_float(), ctypes.c_double()
    for i in range(lib.sb2_profile_count()):
        lib.sb2_profile_get(i, buf, 64, ctypes.byref(ms), ctypes.byref(by))
        tags.add(buf.value.decode())
    lib.sb2_profile_reset()
    return tags


def gpu_hbm(pairs, width, height, levels, use_chroma=0, ref_index=0, xbsep=8, ybsep=8, level0_range=3):
    """pairs: list of (src_planes, ref_planes); all pairs run in one slab / one launch per level."""
    from schroedinger_b200 import device as dev
    count = len(pairs)
    ext = max(xbsep, ybsep)
    ps = dev.Pyramid(width, height, count, levels, ext)
    pr = dev.Pyramid(width, height, count, levels, ext)
    for p, (s, r) in enumerate(pairs):
        for c in range(3):
            ps.slabs[0].upload(p, c, s[c])
            pr.slabs[0].upload(p, c, r[c])
    ps.build()
    pr.build()
    nbx, nby = helpers.hbm_block_counts(width, height, xbsep, ybsep)
    prm = dev.HbmParams(xbsep, ybsep, nbx, nby, ref_index, use_chroma, 1, 1)
    fields = dev.hbm_scan(prm, ps, pr, level0_range)
    torch.cuda.synchronize()
    out = []
    for p in range(count):
        f = np.stack([fl.cpu().numpy().view(helpers.MV_DTYPE).reshape(count, nbx * nby)[p] for fl in fields])
        out.append(f)
    return out, ps


def test_hbm_golden(cuda):
    for idx, (w, h, lv, uc, ri, pan) in enumerate(mg.HBM_GOLDEN_CASES):
        s, r = helpers.panning_pair(w, h, np.random.default_rng(2000 + idx), pan)
        got, ps = gpu_hbm([(s, r)], w, h, lv, uc, ri)
        want = GOLD[f"h{idx}_fields"]
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(got[0][f], want[f]), (idx, f)
        for l in range(lv):
            assert np.array_equal(ps.slabs[l + 1].download(0, 0), GOLD[f"h{idx}_pyr{l}"]), (idx, l)


def test_hbm_many_pairs_one_launch(cuda):
    """GOP-style batch: several independent (frame, ref) pairs side by side."""
    w, h, lv = 320, 192, 3
    pairs = [helpers.panning_pair(w, h, np.random.default_rng(50 + i), (i - 2, 3 - i)) for i in range(5)]
    got, _ = gpu_hbm(pairs, w, h, lv)
    for i, (s, r) in enumerate(pairs):
        want, _, _ = helpers.oracle_hbm(ORACLE, s, r, w, h, levels=lv)
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(got[i][f], want[f]), (i, f)


@pytest.mark.parametrize("case", [(90, 50, 1, 0, 0, (1, 1)), (640, 360, 4, 0, 0, (5, 3)),
                                  (176, 144, 4, 1, 0, (12, 0)), (200, 120, 3, 0, 1, (0, -9)),
                                  (1920, 1080, 4, 0, 0, (5, 3))])
def test_hbm_matches_oracle(cuda, case):
    w, h, lv, uc, ri, pan = case
    s, r = helpers.panning_pair(w, h, np.random.default_rng(w), pan)
    want, _, _ = helpers.oracle_hbm(ORACLE, s, r, w, h, levels=lv, use_chroma=uc, ref_index=ri)
    got, _ = gpu_hbm([(s, r)], w, h, lv, uc, ri)
    for f in ("flags", "metric", "chroma_metric", "v"):
        assert np.array_equal(got[0][f], want[f]), (case, f)


def test_hbm_kernel_selection(cuda, hbm_kernel):
    """The codec's default geometry runs on the wavefront kernel, chroma ME / forced runs on the
    generic one (so the parametrised tests above really cover both)."""
    w, h, lv = 176, 144, 2
    s, r = helpers.panning_pair(w, h, np.random.default_rng(9), (2, 1))
    tags = launched_tags(lambda: gpu_hbm([(s, r)], w, h, lv))
    if hbm_kernel == "wave":
        assert any(t.startswith("hbm_level_s") for t in tags) and "hbm_static" in tags, tags
        assert not any(t.startswith("hbm_generic") for t in tags), tags
        tags = launched_tags(lambda: gpu_hbm([(s, r)], w, h, lv, use_chroma=1))
    assert any(t.startswith("hbm_generic_s") for t in tags) and "hbm_static" not in tags, tags


@pytest.mark.parametrize("case", [(90, 50, 2, (1, 1)), (172, 100, 3, (-3, 2)), (328, 200, 3, (7, -5)),
                                  (1000, 540, 4, (5, 3))])
def test_hbm_incoherent_content(cuda, case):
    """Unrelated noise pictures: every neighbour proposes a different vector, so the ranking of
    left / up / up-left and the de-duplication are exercised on every block; sizes leave partial
    blocks at the right and bottom edges of several levels."""
    w, h, lv, pan = case
    rng = np.random.default_rng(w * 7 + h)
    s, r = helpers.panning_pair(w, h, rng, pan, noise=60)
    r = [rng.integers(0, 256, size=a.shape).astype(np.uint8) if k == 0 else a for k, a in enumerate(r)]
    want, _, _ = helpers.oracle_hbm(ORACLE, s, r, w, h, levels=lv)
    got, _ = gpu_hbm([(s, r)], w, h, lv)
    for f in ("flags", "metric", "chroma_metric", "v"):
        assert np.array_equal(got[0][f], want[f]), (case, f)


@pytest.mark.parametrize("h_range", [1, 2, 3, 4, 5, 7, 10, 13, 20])
def test_hbm_single_level_ranges(cuda, h_range):
    """schro_hierarchical_bm_scan_hint called directly with every class of scan range (the
    wavefront kernel has one instantiation per class) on a level with a parent field."""
    import ctypes
    from schroedinger_b200 import device as dev
    w, h, lv = 232, 136, 2
    s, r = helpers.panning_pair(w, h, np.random.default_rng(77 + h_range), (9, -6), noise=8)
    ps = helpers.build_pyramid(ORACLE, "oracle", s, lv, ext=8)
    pr = helpers.build_pyramid(ORACLE, "oracle", r, lv, ext=8)
    nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
    fn = ORACLE.oracle_hbm_scan_hint
    fn.restype = None
    want = np.zeros((lv + 1, nbx * nby), dtype=helpers.MV_DTYPE)
    for (l, hr) in ((2, 20), (1, h_range)):
        a, b = helpers.pyr_level_struct(ps[l]), helpers.pyr_level_struct(pr[l])
        parent = want[l + 1].ctypes.data_as(ctypes.c_void_p) if l < lv else None
        fn(ctypes.byref(a), ctypes.byref(b), 8, 8, nbx, nby, 0, l, hr, 0, parent,
           want[l].ctypes.data_as(ctypes.c_void_p))
    gs, gr = dev.Pyramid(w, h, 1, lv, 8), dev.Pyramid(w, h, 1, lv, 8)
    for c in range(3):
        gs.slabs[0].upload(0, c, s[c])
        gr.slabs[0].upload(0, c, r[c])
    gs.build()
    gr.build()
    prm = dev.HbmParams(8, 8, nbx, nby, 0, 0, 1, 1)
    f2 = torch.empty(nbx * nby * 20, dtype=torch.uint8, device="cuda")
    f1 = torch.empty(nbx * nby * 20, dtype=torch.uint8, device="cuda")
    dev.hbm_scan_hint(prm, gs.slabs[2], gr.slabs[2], 2, 20, None, f2)
    dev.hbm_scan_hint(prm, gs.slabs[1], gr.slabs[1], 1, h_range, f2, f1)
    torch.cuda.synchronize()
    got = f1.cpu().numpy().view(helpers.MV_DTYPE)
    for f in ("flags", "metric", "chroma_metric", "v"):
        assert np.array_equal(got[f], want[1][f]), (h_range, f)


def test_sad_primitive(cuda):
    import ctypes
    from schroedinger_b200 import lib, check
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, size=(64, 96)).astype(np.uint8)
    b = rng.integers(0, 256, size=(64, 96)).astype(np.uint8)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    for (w, h) in ((8, 8), (12, 12), (16, 7), (32, 9), (5, 3), (1, 1)):
        pos = [(int(rng.integers(0, 96 - w)), int(rng.integers(0, 64 - h)),
                int(rng.integers(0, 96 - w)), int(rng.integers(0, 64 - h))) for _ in range(40)]
        ao = torch.tensor([y * 96 + x for (x, y, _, _) in pos], dtype=torch.int64, device="cuda")
        bo = torch.tensor([y * 96 + x for (_, _, x, y) in pos], dtype=torch.int64, device="cuda")
        out = torch.zeros(len(pos), dtype=torch.int32, device="cuda")
        check(lib.sb2_sad_u8(da.data_ptr(), 96, db.data_ptr(), 96, ao.data_ptr(), bo.data_ptr(), len(pos),
                             w, h, out.data_ptr(), None), "sb2_sad_u8")
        torch.cuda.synchronize()
        want = [int(np.abs(a[y:y + h, x:x + w].astype(int) - b[v:v + h, u:u + w].astype(int)).sum())
                for (x, y, u, v) in pos]
        assert out.cpu().tolist() == want
