"""CPU: the oracle's dequantisation (decoder codeblock loop + Orc programs)
pinned bit-exactly against the compiled, unmodified reference."""
import os

import numpy as np
import pytest

from tests import helpers

ORACLE = helpers.load_oracle()
REF = helpers.load_ref()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (reference tree absent)")
GOLD = os.path.join(helpers.GOLDEN_DIR, "dequant.npz")

CASES = [  # (width, height, depth, hcb, vcb)
    (64, 32, 1, [1, 1], [1, 1]),
    (64, 48, 3, [1, 1, 2, 4], [1, 1, 2, 3]),
    (96, 80, 4, [1, 2, 3, 4, 5], [1, 1, 2, 3, 5]),
    (160, 96, 5, [1, 1, 1, 2, 4, 7], [1, 1, 1, 2, 3, 5]),
    (48, 16, 2, [3, 5, 7], [2, 3, 4]),
]


def make_case(rng, dtype, w, h, depth, hcb, vcb, tables, full):
    if dtype == np.int16:
        a = rng.integers(-32768, 32768, size=(h, w)) if full else rng.integers(-40, 41, size=(h, w))
    else:
        a = rng.integers(-2 ** 31, 2 ** 31, size=(h, w)) if full else rng.integers(-4000, 4001, size=(h, w))
    a = a.astype(dtype)
    a[rng.random(a.shape) < 0.4] = 0                      # most quantised coefficients are zero
    n = helpers.dequant_table_size(depth, hcb, vcb)
    idx = rng.integers(0, 61, size=n)
    offs = tables[1] if rng.random() < 0.5 else tables[2]
    quant = np.stack([tables[0][idx].astype(np.int64), offs[idx].astype(np.int64) + 2], axis=1).astype(np.int32)
    return a, quant


@needs_ref
@pytest.mark.parametrize("dtype", [np.int16, np.int32])
def test_dequantise_matches_reference(dtype):
    rng = np.random.default_rng(5)
    tables = helpers.ref_quant_tables(REF)
    for (w, h, depth, hcb, vcb) in CASES:
        for full in (False, True):
            a, quant = make_case(rng, dtype, w, h, depth, hcb, vcb, tables, full)
            want = helpers.cpu_dequantise(REF, "ref", a, depth, hcb, vcb, quant)
            got = helpers.cpu_dequantise(ORACLE, "oracle", a, depth, hcb, vcb, quant)
            assert np.array_equal(got, want), (dtype, w, h, depth, full)
            assert not np.array_equal(got, a)


def test_dequantise_golden():
    """tests/golden/dequant.npz: inputs, quantiser pairs (from the reference's tables) and the
    compiled reference's outputs; readable where the reference is absent."""
    g = np.load(GOLD)
    n = int(g["ncases"])
    assert n >= 10
    for i in range(n):
        depth, hcb, vcb = int(g[f"c{i}_depth"]), g[f"c{i}_hcb"].tolist(), g[f"c{i}_vcb"].tolist()
        got = helpers.cpu_dequantise(ORACLE, "oracle", g[f"c{i}_in"], depth, hcb, vcb, g[f"c{i}_quant"])
        assert np.array_equal(got, g[f"c{i}_out"]), i
