"""CPU: the wavelet oracle against the committed golden vectors (made from the unmodified
reference) and, when oracle/_ref is present, against the reference itself on the
reference's own sweep (testsuite/wavelet_2d.c:282-299: 66 patterns at 20x20, every even
size 2..40 x 2..40 on random data, s16 and s32)."""
import os

import numpy as np
import pytest

from tests import helpers

ORACLE = helpers.load_oracle()
REF = helpers.load_ref()
GOLD = np.load(os.path.join(helpers.GOLDEN_DIR, "wavelet.npz"))


def test_oracle_matches_golden_vectors():
    n = 0
    for key in GOLD.files:
        if not key.endswith("_in"):
            continue
        base = key[:-3]
        filt = int(base.split("_")[1][1:])
        depth = 3 if base.endswith("ml3") else None
        for d in ("fwd", "inv"):
            got = helpers.cpu_wavelet(ORACLE, "oracle", d, GOLD[key].copy(), filt, depth)
            assert np.array_equal(got, GOLD[f"{base}_{d}"]), (base, d)
            n += 1
    assert n >= 250


@pytest.mark.parametrize("dtype", [np.int16, np.int32])
@pytest.mark.parametrize("filt", range(7))
def test_round_trip_is_exact(filt, dtype):
    """forward then inverse restores the input (values small enough not to wrap)."""
    rng = np.random.default_rng(filt)
    a = rng.integers(-255, 256, size=(64, 96)).astype(dtype)
    b = helpers.cpu_wavelet(ORACLE, "oracle", "fwd", a.copy(), filt, 3)
    c = helpers.cpu_wavelet(ORACLE, "oracle", "inv", b, filt, 3)
    assert np.array_equal(a, c)


@pytest.mark.skipif(REF is None, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dtype", [np.int16, np.int32])
@pytest.mark.parametrize("filt", range(7))
def test_oracle_matches_reference_sweep(filt, dtype):
    rng = np.random.default_rng(100 + filt)
    amp_full = 32767 if dtype == np.int16 else 2 ** 31 - 1
    cases = []
    for name, p in helpers.patterns(20, 20, dtype, rng):
        cases.append(p)
    for h in range(2, 41, 2):
        for w in range(2, 41, 6):
            cases.append(rng.integers(-255, 256, size=(h, w)).astype(dtype))
    for w in range(2, 41, 2):
        cases.append(rng.integers(-amp_full, amp_full + 1, size=(12, w)).astype(dtype))
    for a in cases:
        for d in ("fwd", "inv"):
            r = helpers.cpu_wavelet(REF, "ref", d, a.copy(), filt)
            o = helpers.cpu_wavelet(ORACLE, "oracle", d, a.copy(), filt)
            assert np.array_equal(r, o), (filt, dtype, a.shape, d)


@pytest.mark.skipif(REF is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_reference_multilevel_1080p():
    rng = np.random.default_rng(7)
    for dtype, filt, depth, shape in ((np.int16, 1, 4, (1088, 1920)), (np.int16, 0, 4, (544, 960)),
                                      (np.int32, 6, 5, (1088, 1920))):
        a = rng.integers(-512, 512, size=shape).astype(dtype)
        for d in ("fwd", "inv"):
            r = helpers.cpu_wavelet(REF, "ref", d, a.copy(), filt, depth)
            o = helpers.cpu_wavelet(ORACLE, "oracle", d, a.copy(), filt, depth)
            assert np.array_equal(r, o)
