"""oracle_split2.c pinned against the compiled reference's schro_do_split2 + schro_motion_copy_to
(schroedinger/schromotionest.c:1601-1802, 1511-1523; reached through oracle/ref_me_static.c, which compiles
the reference's schromotionest.c unmodified with one entry point appended).  CPU only."""
import numpy as np
import pytest

from tests import helpers

ref_me = helpers.load_ref_me()
oracle = helpers.load_oracle()
needs_ref = pytest.mark.skipif(ref_me is None, reason="oracle/_ref not built")


def _golden_check():
    import os
    from tests.golden import make_golden as mg
    gold = np.load(os.path.join(helpers.GOLDEN_DIR, "split2.npz"))
    for idx, case in enumerate(mg.SPLIT2_GOLDEN_CASES):
        w, h, prec, nrefs, lam, seed = case
        src, refs, _ = mg.split2_inputs(oracle, idx, case)
        fields = [gold[f"s{idx}_field{r}"] for r in range(nrefs)]
        got = helpers.oracle_split2(oracle, src, refs, fields, w, h, 8, 8, prec, lam)
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(got[0][f], gold[f"s{idx}_motion"][f]), (idx, f)
        assert np.array_equal(got[1], gold[f"s{idx}_sb_error"]) and np.array_equal(got[2], gold[f"s{idx}_sb_entropy"]), idx


def _check(w, h, prec, lam, seed, num_refs=2, pans=((5, 3), (-4, 2)), mutate=None):
    rng = np.random.default_rng(seed)
    src, refs, fields = helpers.split2_case(oracle, w, h, rng, prec, num_refs, pans, lam)
    if mutate:
        mutate(src, refs, fields, rng)
    want = helpers.ref_split2(ref_me, src, refs, fields, w, h, 8, 8, prec, lam)
    got = helpers.oracle_split2(oracle, src, refs, fields, w, h, 8, 8, prec, lam)
    for f in ("flags", "metric", "chroma_metric", "v"):
        bad = np.flatnonzero(np.any(np.atleast_2d(got[0][f] != want[0][f]).reshape(len(got[0]), -1), axis=1))
        assert bad.size == 0, (f, bad[:8], got[0][bad[:4]], want[0][bad[:4]])
    assert np.array_equal(got[1], want[1])
    assert np.array_equal(got[2], want[2])
    modes = got[0]["flags"] & 3
    return [int((modes == m).sum()) for m in range(4)]


@needs_ref
@pytest.mark.parametrize("prec", [0, 1, 2, 3])
@pytest.mark.parametrize("lam", [0.0, 0.1, 2.0])
def test_split2_matches_reference(prec, lam):
    _check(176, 144, prec, lam, seed=prec * 7 + int(lam))


@needs_ref
def test_split2_mode_mix():
    """A picture whose halves favour different references, plus flat patches that go DC."""
    def mutate(src, refs, fields, rng):
        h, w = src[0].shape
        refs[0][0][:, w // 2:] = rng.integers(0, 256, size=(h, w - w // 2))           # ref 0 useless on the right
        refs[1][0][:, :w // 3] = rng.integers(0, 256, size=(h, w // 3))               # ref 1 useless on the left
        for k in range(3):
            hh, ww = src[k].shape
            src[k][hh // 2:, ww // 4:ww // 2] = 90 + 10 * k                              # flat patch: nothing matches it
        for r, f in enumerate(fields):                                                  # luma metrics that go with it
            f["metric"] = rng.integers(0, 4000, size=len(f))
    counts = _check(192, 160, 2, 0.3, seed=11, mutate=mutate)
    assert all(c > 0 for c in counts), counts                                           # DC, ref 0, ref 1 and biref all occur


@needs_ref
def test_split2_single_reference_and_ragged():
    _check(200, 104, 2, 0.25, seed=3, num_refs=1)
    _check(100, 70, 3, 0.05, seed=4)
    _check(100, 70, 1, 0.5, seed=5, num_refs=1)


@needs_ref
def test_split2_biref_range_test():
    """Vectors near the edge of the extended frame: the bi-reference candidate is dropped (:1719-1724)."""
    def mutate(src, refs, fields, rng):
        for r, f in enumerate(fields):
            n = len(f)
            idx = rng.choice(n, n // 4, replace=False)
            f["v"][idx, r] = rng.integers(-30, 31, size=len(idx))
            f["v"][idx, 2 + r] = rng.integers(-30, 31, size=len(idx))
    _check(176, 144, 1, 0.1, seed=21, mutate=mutate)
    _check(176, 144, 0, 0.1, seed=22, mutate=mutate)


def test_oracle_vs_golden_fixture():
    """tests/golden/split2.npz holds outputs of the compiled reference (make_golden.py: split2_golden); this
    comparison also runs where oracle/_ref is absent."""
    _golden_check()


def _synthetic_fields(rng, w, h, bs, prec, nrefs):
    nbx, nby = helpers.hbm_block_counts(w, h, bs, bs)
    fields = [np.zeros(nbx * nby, helpers.MV_DTYPE) for _ in range(nrefs)]
    for r, f in enumerate(fields):
        f["flags"] = r + 1
        f["v"][:, r] = rng.integers(-6 << prec, (6 << prec) + 1, size=nbx * nby)
        f["v"][:, 2 + r] = rng.integers(-6 << prec, (6 << prec) + 1, size=nbx * nby)
        f["metric"] = rng.integers(0, 3000, size=nbx * nby)
    return fields


@needs_ref
@pytest.mark.parametrize("w,h,bs,prec", [(192, 144, 12, 2), (256, 128, 16, 3), (96, 64, 4, 1), (200, 104, 12, 0)])
def test_split2_other_block_sizes(w, h, bs, prec):
    rng = np.random.default_rng(w + bs + prec)
    src, refs, _ = helpers.subpel_case(oracle, w, h, rng)
    fields = _synthetic_fields(rng, w, h, bs, prec, 2)
    want = helpers.ref_split2(ref_me, src, refs, fields, w, h, bs, bs, prec, 0.2)
    got = helpers.oracle_split2(oracle, src, refs, fields, w, h, bs, bs, prec, 0.2)
    for f in ("flags", "metric", "chroma_metric", "v"):
        assert np.array_equal(got[0][f], want[0][f]), f
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])


@needs_ref
@pytest.mark.parametrize("seed", range(8))
def test_split2_random_sweep(seed):
    """random size (ragged or not), precision, number of references, lambda, block size and field mutations"""
    rng = np.random.default_rng(1000 + seed)
    bs = int(rng.choice([4, 8, 8, 12, 16]))
    w = int(rng.integers(5, 14)) * 16 + int(rng.choice([0, 0, 2, 6, 10]))
    h = int(rng.integers(4, 10)) * 16 + int(rng.choice([0, 0, 2, 6, 10]))
    prec, nrefs, lam = int(rng.integers(0, 4)), int(rng.integers(1, 3)), float(rng.choice([0.0, 0.03, 0.3, 3.0]))
    src, refs, _ = helpers.subpel_case(oracle, w, h, rng, num_refs=nrefs)
    fields = _synthetic_fields(rng, w, h, bs, prec, nrefs)
    if seed & 1:                                   # flat patches: DC blocks
        for k in range(3):
            hh, ww = src[k].shape
            src[k][hh // 3:2 * hh // 3, ww // 4:3 * ww // 4] = 60 + 30 * k
    want = helpers.ref_split2(ref_me, src, refs, fields, w, h, bs, bs, prec, lam)
    got = helpers.oracle_split2(oracle, src, refs, fields, w, h, bs, bs, prec, lam)
    for f in ("flags", "metric", "chroma_metric", "v"):
        assert np.array_equal(got[0][f], want[0][f]), (seed, bs, w, h, prec, nrefs, lam, f)
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
