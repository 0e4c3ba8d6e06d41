"""CPU: the oracle's combine / convert glue (schro_frame_convert, schro_frame_add / _subtract;
SURVEY.md 8f rank 2) pinned bit-exactly against the compiled, unmodified reference."""
import numpy as np
import pytest

from tests import helpers

ORACLE = helpers.load_oracle()
REF = helpers.load_ref()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (reference tree absent)")

SIZES = [((64, 48), (64, 48)), ((70, 50), (64, 48)), ((64, 48), (72, 56)), ((33, 17), (40, 24)), ((2, 2), (2, 2))]


@needs_ref
@pytest.mark.parametrize("sdepth,ddepth", [(0, 1), (0, 2), (1, 0), (2, 0), (1, 2), (2, 1), (0, 0), (1, 1), (2, 2)])
def test_convert_matches_reference(sdepth, ddepth):
    rng = np.random.default_rng(10 * sdepth + ddepth)
    for (sw, sh), (dw, dh) in SIZES:
        for full in (True, False):
            src = helpers.random_planes(rng, sdepth, sw, sh, full)
            want = helpers.ref_convert(REF, src, sdepth, sw, sh, ddepth, dw, dh)
            got = helpers.oracle_convert(ORACLE, src, sdepth, sw, sh, ddepth, dw, dh)
            for k in range(3):
                assert np.array_equal(got[k], want[k]), (sdepth, ddepth, (sw, sh), (dw, dh), full, k)


@needs_ref
@pytest.mark.parametrize("sdepth", [0, 1])
@pytest.mark.parametrize("subtract", [0, 1])
def test_add_subtract_match_reference(sdepth, subtract):
    rng = np.random.default_rng(3 + sdepth)
    for (sw, sh), (dw, dh) in SIZES:
        dst = helpers.random_planes(rng, 1, dw, dh, True)
        src = helpers.random_planes(rng, sdepth, sw, sh, True)
        want = helpers.cpu_add(REF, "ref", dst, dw, dh, src, sdepth, sw, sh, subtract)
        got = helpers.cpu_add(ORACLE, "oracle", dst, dw, dh, src, sdepth, sw, sh, subtract)
        for k in range(3):
            assert np.array_equal(got[k], want[k]), (sdepth, subtract, (sw, sh), (dw, dh), k)


def test_convert_known_answers():
    """The wrap / saturation points of the Orc programs, spelled out (schroorc.orc:476-549)."""
    s16 = [np.array([[-32768, -129, -128, 0, 127, 128, 32639, 32640, 32767]], np.int16)] + [np.zeros((1, 5), np.int16)] * 2
    got = helpers.oracle_convert(ORACLE, s16, 1, 9, 1, 0, 9, 1)[0][0]
    # s + 128 wraps at 16 bits (addw) before the saturating narrow: 32640 + 128 -> -32768 -> 0
    assert got.tolist() == [0, 0, 0, 128, 255, 255, 255, 0, 0]
    s32 = [np.array([[-2 ** 31, -129, -128, 127, 128, 32639, 32640, 2 ** 31 - 129, 2 ** 31 - 128, 2 ** 31 - 1]], np.int32)] + [np.zeros((1, 5), np.int32)] * 2
    got = helpers.oracle_convert(ORACLE, s32, 2, 10, 1, 0, 10, 1)[0][0]
    # s + 128 wraps at 32 bits (addl), saturates to 0..65535 (convsuslw), and that word is read as
    # signed by convsuswb: 32640 + 128 = 32768 -> -32768 -> 0 (schroorc-dist.c:4274-4306)
    assert got.tolist() == [0, 0, 0, 255, 255, 255, 0, 0, 0, 0]
    u8 = [np.array([[0, 127, 128, 255]], np.uint8)] + [np.zeros((1, 2), np.uint8)] * 2
    assert helpers.oracle_convert(ORACLE, u8, 0, 4, 1, 1, 4, 1)[0][0].tolist() == [-128, -1, 0, 127]
    assert helpers.oracle_convert(ORACLE, u8, 0, 4, 1, 2, 6, 2)[0].tolist() == [[-128, -1, 0, 127, 127, 127]] * 2
    t = [np.array([[70000, -70000, 32768]], np.int32)] + [np.zeros((1, 2), np.int32)] * 2
    assert helpers.oracle_convert(ORACLE, t, 2, 3, 1, 1, 3, 1)[0][0].tolist() == [4464, -4464, -32768]   # convlw truncates


def test_golden_fixtures():
    """tests/golden/glue.npz: outputs of the compiled reference (made by tests/golden/make_golden.py),
    readable where the reference itself is absent."""
    import os
    import re
    g = np.load(os.path.join(helpers.GOLDEN_DIR, "glue.npz"))
    nconv = nadd = 0
    for key in sorted({k.rsplit("_", 1)[0] for k in g.files}):
        m = re.match(r"conv_(\d)(\d)_(\d+)x(\d+)_(\d+)x(\d+)$", key)
        if m:
            sd, dd, sw, sh, dw, dh = map(int, m.groups())
            src = [g[f"{key}_in{c}"] for c in range(3)]
            got = helpers.oracle_convert(ORACLE, src, sd, sw, sh, dd, dw, dh)
            for c in range(3):
                assert np.array_equal(got[c], g[f"{key}_out{c}"]), (key, c)
            nconv += 1
            continue
        m = re.match(r"add_(\d)_(\d)_(\d+)x(\d+)_(\d+)x(\d+)$", key)
        if m:
            sd, sub, sw, sh, dw, dh = map(int, m.groups())
            dst = [g[f"{key}_dst{c}"] for c in range(3)]
            src = [g[f"{key}_src{c}"] for c in range(3)]
            got = helpers.cpu_add(ORACLE, "oracle", dst, dw, dh, src, sd, sw, sh, sub)
            for c in range(3):
                assert np.array_equal(got[c], g[f"{key}_out{c}"]), (key, c)
            nadd += 1
    assert nconv == 18 and nadd == 12, (nconv, nadd)
