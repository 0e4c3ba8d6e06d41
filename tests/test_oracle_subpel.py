"""oracle_subpel.c pinned against the compiled reference's schro_encoder_motion_predict_subpel_deep
(schroedinger/schromotionest.c:246-355), driven through a SchroMe the reference builds itself.  CPU only."""
import numpy as np
import pytest

from tests import helpers

ref = helpers.load_ref()
oracle = helpers.load_oracle()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built")


def _check(w, h, prec, lam, seed, num_refs=2, pans=((5, 3), (-4, 2)), bs=8, mutate=None):
    rng = np.random.default_rng(seed)
    src, refs, fields = helpers.subpel_case(oracle, w, h, rng, pans=pans, num_refs=num_refs)
    if bs != 8:
        nbx, nby = helpers.hbm_block_counts(w, h, bs, bs)
        fields = [np.zeros(nbx * nby, helpers.MV_DTYPE) for _ in refs]
        for r, f in enumerate(fields):
            f["flags"] = r + 1
            f["v"][:, r] = rng.integers(-6, 7, size=nbx * nby)
            f["v"][:, 2 + r] = rng.integers(-6, 7, size=nbx * nby)
            f["metric"] = rng.integers(0, 3000, size=nbx * nby)
    if mutate:
        mutate(fields, rng)
    want = helpers.ref_subpel(ref, src, refs, fields, w, h, bs, bs, prec, lam)
    got = helpers.oracle_subpel(oracle, src, refs, fields, w, h, bs, bs, prec, lam)
    moved = 0
    for r in range(len(refs)):
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(got[r][f], want[r][f]), (r, f)
        moved += int(np.any(got[r]["v"] != (fields[r]["v"].astype(np.int32) << prec), axis=1).sum())
    return moved


@pytest.mark.parametrize("prec", [1, 2, 3])
@pytest.mark.parametrize("lam", [0.0, 0.1, 1.0, 10.0])
def test_subpel_matches_reference(prec, lam):
    def jitter(fields, rng):
        # integer pans are matched exactly by the block matcher; knock a third of the vectors off by one
        for r, f in enumerate(fields):
            idx = rng.choice(len(f), len(f) // 3, replace=False)
            f["v"][idx, r] += rng.integers(-1, 2, size=len(idx)).astype(np.int16)
            f["v"][idx, 2 + r] += rng.integers(-1, 2, size=len(idx)).astype(np.int16)
    moved = _check(176, 144, prec, lam, seed=prec * 10 + int(lam), mutate=jitter)
    assert moved > 0                     # the refinement did move vectors


def test_subpel_single_reference_and_ragged():
    _check(200, 104, 2, 0.25, seed=3, num_refs=1)
    _check(100, 70, 3, 0.05, seed=4)


def test_subpel_other_block_sizes():
    _check(192, 144, 2, 0.1, seed=5, bs=12)
    _check(256, 128, 2, 0.1, seed=6, bs=16)


def test_subpel_vectors_at_the_range_limits():
    """Vectors that point to the edge of the extended reference: probes fail the range test
    (schromotionest.c:306-312) one by one."""
    def mutate(fields, rng):
        for r, f in enumerate(fields):
            n = len(f)
            idx = rng.choice(n, n // 3, replace=False)
            f["v"][idx, r] = rng.integers(-40, 41, size=len(idx))
            f["v"][idx, 2 + r] = rng.integers(-40, 41, size=len(idx))
    # keep the start vectors themselves inside the reference (pixel range +-8 around the picture)
    def mutate_safe(fields, rng):
        mutate(fields, rng)
        for r, f in enumerate(fields):
            f["v"][:, r] = np.clip(f["v"][:, r], -6, 6)
            f["v"][:, 2 + r] = np.clip(f["v"][:, 2 + r], -6, 6)
    _check(128, 96, 2, 0.1, seed=8, mutate=mutate_safe)


@pytest.mark.parametrize("seed", range(6))
def test_subpel_random_sweep(seed):
    """random size, precision, lambda, block size, number of references"""
    rng = np.random.default_rng(500 + seed)
    bs = int(rng.choice([8, 8, 12, 16]))
    w = int(rng.integers(6, 14)) * 16 + int(rng.choice([0, 0, 4, 10]))
    h = int(rng.integers(4, 10)) * 16 + int(rng.choice([0, 0, 6, 12]))
    prec, lam, nrefs = int(rng.integers(1, 4)), float(rng.choice([0.0, 0.02, 0.5, 4.0])), int(rng.integers(1, 3))
    _check(w, h, prec, lam, seed=600 + seed, num_refs=nrefs, bs=bs)
