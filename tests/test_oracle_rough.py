"""oracle_rough.c pinned against the compiled reference's schro_rough_me_heirarchical_scan
(schroedinger/schroroughmotion.c:46-300).  CPU only."""
import numpy as np
import pytest

from tests import helpers

ref = helpers.load_ref()
oracle = helpers.load_oracle()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built")


def _compare(w, h, levels, pan, seed, xbsep=8, ybsep=8, dists=(12, 4), ref_index=0, noise=3, strict=True):
    rng = np.random.default_rng(seed)
    s, r = helpers.panning_pair(w, h, rng, pan, noise=noise)
    want = helpers.ref_rough(ref, s, r, w, h, xbsep, ybsep, levels, ref_index, *dists)
    got, _, _ = helpers.oracle_rough(oracle, s, r, w, h, xbsep, ybsep, levels, ref_index, *dists)
    nbx, nby = helpers.hbm_block_counts(w, h, xbsep, ybsep)
    for l in range(levels, 0, -1):
        inside = np.ones(nbx * nby, bool) if strict else helpers.rough_inside_mask(w, h, xbsep, ybsep, nbx, nby, l)
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(got[l][f][inside], want[l][f][inside]), (l, f)
    assert not got[0]["metric"].any()
    return got


@pytest.mark.parametrize("w,h,levels,pan", [(256, 128, 2, (5, 3)), (512, 256, 3, (-7, 2)), (1024, 512, 4, (11, -6))])
def test_rough_matches_reference(w, h, levels, pan):
    got = _compare(w, h, levels, pan, seed=w)
    # the pan is recovered by most blocks of the finest rough level
    nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
    grid = got[1]["v"].reshape(nby, nbx, 4)[::2, ::2].astype(int)
    hit = np.mean((np.abs(grid[..., 0] + pan[0]) <= 2) & (np.abs(grid[..., 2] + pan[1]) <= 2))
    assert levels > 3 or hit > 0.8, hit      # (the 4-level case locks onto the texture's period)


def test_rough_second_reference_and_other_distances():
    _compare(256, 192, 3, (4, -4), seed=5, dists=(7, 2), ref_index=1)
    _compare(256, 192, 2, (20, 9), seed=6, dists=(20, 6))


def test_rough_other_block_sizes():
    _compare(384, 288, 2, (3, 1), seed=7, xbsep=12, ybsep=12)
    _compare(256, 256, 2, (-2, 6), seed=8, xbsep=16, ybsep=8)


def test_rough_noise_only():
    # incoherent content: candidates disagree everywhere, ties are frequent on flat pictures
    rng = np.random.default_rng(9)
    w, h, levels = 256, 128, 3
    flat = [np.full((h, w), 77, np.uint8), np.full((h // 2, w // 2), 10, np.uint8), np.full((h // 2, w // 2), 200, np.uint8)]
    noisy = [rng.integers(0, 256, size=a.shape).astype(np.uint8) for a in flat]
    for (s, r) in ((flat, flat), (noisy, flat), (noisy, [a[::-1].copy() for a in noisy])):
        want = helpers.ref_rough(ref, s, r, w, h, levels=levels)
        got, _, _ = helpers.oracle_rough(oracle, s, r, w, h, levels=levels)
        for l in range(levels, 0, -1):
            for f in ("flags", "metric", "v"):
                assert np.array_equal(got[l][f], want[l][f]), (l, f)


@pytest.mark.parametrize("w,h,levels", [(176, 144, 2), (360, 270, 3), (200, 72, 3)])
def test_rough_ragged_sizes(w, h, levels):
    """Sizes whose block grid over-covers the picture: blocks overlapping the frame must match; the
    blocks entirely outside are where the reference reads stale memory (oracle_rough.c)."""
    _compare(w, h, levels, (3, -2), seed=w + h, strict=False)
