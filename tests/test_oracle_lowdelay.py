"""oracle_lowdelay.c pinned against the compiled reference's schro_decoder_decode_lowdelay_transform_data
(schroedinger/schrolowdelay.c:99-761): the dispatcher, and its _slow and _fast paths on their own.  The
slices are packed by tests/helpers.lowdelay_encode.  CPU only."""
import numpy as np
import pytest

from tests import helpers

ref = helpers.load_ref()
oracle = helpers.load_oracle()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built")


def quantised_planes(rng, w, h, amp, dtype=np.int32):
    # mostly small values with a long tail, like a real quantised picture
    def one(hh, ww):
        a = rng.geometric(0.45, size=(hh, ww)) - 1
        a = a * rng.choice([-1, 1], size=(hh, ww))
        big = rng.random((hh, ww)) < 0.02
        a[big] = rng.integers(-amp, amp + 1, size=int(big.sum()))
        return a.astype(dtype)
    return [one(h, w), one(h // 2, w // 2), one(h // 2, w // 2)]


CASES = [
    # w, h, depth, n_horiz, n_vert, slice_bytes num / denom, is_s32
    (64, 32, 2, 4, 2, 97, 2, 0),          # fractional slice size, chroma LL 8x4 over 4x2 slices: the fast path
    (96, 64, 3, 3, 4, 640, 3, 0),         # slices that do not divide the LL band: the slow path
    (128, 64, 3, 8, 4, 40, 1, 0),
    (128, 64, 3, 8, 4, 40, 1, 1),         # s32: always the slow_s32 path
    (96, 96, 2, 6, 6, 33, 1, 1),
    (480, 288, 4, 15, 9, 75, 2, 0),
]


def run_case(case, seed, truncate=0.0, amp=200):
    w, h, depth, nh, nv, num, denom, is_s32 = case
    rng = np.random.default_rng(seed)
    tables = helpers.ref_quant_tables(ref)
    qm = [int(v) for v in rng.integers(0, 8, size=1 + 3 * depth)]
    q = quantised_planes(rng, w, h, amp)
    aligned = ((w // 2) >> depth) % nh == 0 and ((h // 2) >> depth) % nv == 0
    fast = (not is_s32) and aligned
    data, bases = helpers.lowdelay_encode(q, depth, nh, nv, num, denom, rng, truncate=truncate, fast_lengths=fast)
    want = helpers.cpu_lowdelay(ref, "ref", data, w, h, depth, nh, nv, num, denom, qm, is_s32, 0)
    got = helpers.cpu_lowdelay(oracle, "oracle", data, w, h, depth, nh, nv, num, denom, qm, is_s32, 1 if fast else 0, tables)
    for c in range(3):
        assert np.array_equal(got[c], want[c]), (case, c, "dispatcher")
    if not is_s32:
        # the other s16 path on the same bytes (the length field is read with the width that path uses)
        other = 1 if fast else (2 if aligned else None)
        if other is not None and helpers.ilog2up(8 * (num // denom)) == helpers.ilog2up(8 * (num // denom + 1)):
            want2 = helpers.cpu_lowdelay(ref, "ref", data, w, h, depth, nh, nv, num, denom, qm, 0, other)
            got2 = helpers.cpu_lowdelay(oracle, "oracle", data, w, h, depth, nh, nv, num, denom, qm, 0, 1 if other == 2 else 0, tables)
            for c in range(3):
                assert np.array_equal(got2[c], want2[c]), (case, c, "other path")
    return got, q, bases


@pytest.mark.parametrize("case", CASES)
def test_lowdelay_matches_reference(case):
    got, q, bases = run_case(case, seed=sum(case))
    assert any(np.any(p != 0) for p in got)


@pytest.mark.parametrize("case", CASES[:5])
def test_lowdelay_truncated_and_overflowing_slices(case):
    """Luma lengths cut short and slices too small for their coefficients: the readers run into the 1-bit guard."""
    w, h, depth, nh, nv, num, denom, is_s32 = case
    run_case(case, seed=7 + sum(case), truncate=0.5)
    run_case((w, h, depth, nh, nv, max(4 * denom, num // 3), denom, is_s32), seed=8 + sum(case))


def test_lowdelay_large_values_wrap():
    """Quantised values whose products leave 16 bits: the fast path's Orc arithmetic wraps, the slow path truncates."""
    for case in (CASES[0], CASES[1]):
        run_case(case[:5] + (case[5] * 4, case[6], case[7]), seed=3, amp=6000)


def test_lowdelay_random_bytes():
    """Arbitrary bytes are a valid stream: every slice decodes to something; limited to base indices that keep
    exp-Golomb run lengths representable (random bytes hold no long zero runs)."""
    rng = np.random.default_rng(11)
    tables = helpers.ref_quant_tables(ref)
    for (w, h, depth, nh, nv, num, denom, is_s32) in CASES[:5]:
        total = sum(helpers.lowdelay_slice_sizes(num, denom, nh, nv))
        raw = rng.integers(0, 256, size=total, dtype=np.uint8)
        aligned = ((w // 2) >> depth) % nh == 0 and ((h // 2) >> depth) % nv == 0
        fast = (not is_s32) and aligned
        pos = 0
        for sz in helpers.lowdelay_slice_sizes(num, denom, nh, nv):
            bits = np.unpackbits(raw[pos:pos + sz])
            # base indices below 60: the reference's fast path has no tables beyond (schrolowdelay.c:473)
            base = int(rng.integers(0, 60))
            bits[:7] = [(base >> (6 - i)) & 1 for i in range(7)]
            # a declared luma length that points past the END OF THE BUFFER makes the reference read its heap:
            # such slices (the last one or two) get a length that stays inside; lengths that run into the
            # FOLLOWING slices stay as they are -- the reference reads on, and so must the oracle
            lb = helpers.ilog2up(8 * ((num // denom) if fast else sz))
            ylen = int("".join(map(str, bits[7:7 + lb])), 2)
            left = 8 * (total - pos) - 7 - lb
            if ylen > left:
                ylen = int(rng.integers(0, left + 1))
                bits[7:7 + lb] = [(ylen >> (lb - 1 - i)) & 1 for i in range(lb)]
            raw[pos:pos + sz] = np.packbits(bits)
            pos += sz
        data = bytes(raw)
        qm = [0] * (1 + 3 * depth)
        want = helpers.cpu_lowdelay(ref, "ref", data, w, h, depth, nh, nv, num, denom, qm, is_s32, 0)
        got = helpers.cpu_lowdelay(oracle, "oracle", data, w, h, depth, nh, nv, num, denom, qm, is_s32, 1 if fast else 0, tables)
        for c in range(3):
            assert np.array_equal(got[c], want[c]), ((w, h), c)
