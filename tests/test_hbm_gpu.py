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
