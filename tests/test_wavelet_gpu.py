"""GPU parity: the CUDA wavelets (through the sb2_* C-ABI) against the oracle, bit-exact."""
import os

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

ORACLE = helpers.load_oracle()
GOLD = np.load(os.path.join(helpers.GOLDEN_DIR, "wavelet.npz"))


def gpu_iwt(direction, arrays, filt, depth, in_place=False):
    """arrays: list of same-dtype 2-D planes = components of ONE picture."""
    from schroedinger_b200 import device as dev
    depth_name = "s32" if arrays[0].dtype == np.int32 else "s16"
    layout = dev.FrameLayout(depth_name, [(a.shape[1], a.shape[0]) for a in arrays])
    src = dev.PictureSlab(layout, 1)
    dst = src if in_place else dev.PictureSlab(layout, 1)
    for c, a in enumerate(arrays):
        src.upload(0, c, a)
    fn = dev.iwt_forward if direction == "fwd" else dev.iwt_inverse
    fn(src, dst, filt, depth)
    return [dst.download(0, c) for c in range(len(arrays))]


def test_golden_vectors(cuda):
    n = 0
    for key in GOLD.files:
        if not key.endswith("_in"):
            continue
        base = key[:-3]
        filt = int(base.split("_")[1][1:])
        depth = 3 if base.endswith("ml3") else 1
        for d in ("fwd", "inv"):
            got = gpu_iwt(d, [GOLD[key]], filt, depth)[0]
            assert np.array_equal(got, GOLD[f"{base}_{d}"]), (base, d)
            n += 1
    assert n >= 250


@pytest.mark.parametrize("dtype", [np.int16, np.int32])
@pytest.mark.parametrize("filt", range(7))
def test_reference_sweep_sizes_and_patterns(cuda, filt, dtype):
    """The reference's own sweep (testsuite/wavelet_2d.c:282-299) + full-range values."""
    rng = np.random.default_rng(200 + filt)
    amp_full = 32767 if dtype == np.int16 else 2 ** 31 - 1
    cases = [p for _, p in helpers.patterns(20, 20, dtype, rng)]
    for h in range(2, 41, 2):
        for w in range(2, 41, 10):
            cases.append(rng.integers(-255, 256, size=(h, w)).astype(dtype))
    for w in range(2, 41, 4):
        cases.append(rng.integers(-amp_full, amp_full + 1, size=(14, w)).astype(dtype))
    for a in cases:
        for d in ("fwd", "inv"):
            want = helpers.cpu_wavelet(ORACLE, "oracle", d, a.copy(), filt)
            got = gpu_iwt(d, [a], filt, 1)[0]
            assert np.array_equal(got, want), (filt, dtype, a.shape, d)


@pytest.mark.parametrize("case", [
    (np.int16, 1, 4, (1088, 1920), (544, 960)),      # config 1: LeGall 4 levels 1080p 4:2:0
    (np.int16, 0, 4, (1088, 1920), (544, 960)),      # config 2: DD 9/7 4 levels
    (np.int32, 6, 5, (2176, 3840), (1088, 1920)),    # config 3: Daubechies 5 levels 2160p s32
    (np.int16, 2, 3, (264, 520), (136, 264)),
    (np.int32, 5, 2, (132, 260), (68, 132)),
    (np.int16, 5, 3, (200, 328), (104, 168)),
    (np.int32, 3, 3, (72, 200), (40, 104)),
    (np.int16, 4, 2, (36, 100), (20, 52)),
])
def test_multilevel_frames_match_oracle(cuda, case):
    dtype, filt, depth, luma, chroma = case
    rng = np.random.default_rng(filt * 10 + depth)
    planes = [rng.integers(-600, 600, size=s).astype(dtype) for s in (luma, chroma, chroma)]
    for d in ("fwd", "inv"):
        want = [helpers.cpu_wavelet(ORACLE, "oracle", d, p.copy(), filt, depth) for p in planes]
        got = gpu_iwt(d, planes, filt, depth)
        for c in range(3):
            assert np.array_equal(got[c], want[c]), (case, d, c)
        got = gpu_iwt(d, planes, filt, depth, in_place=True)
        for c in range(3):
            assert np.array_equal(got[c], want[c]), (case, d, c, "in place")


def test_batch_of_pictures_and_round_trip(cuda):
    """A slab of pictures in one launch; forward then inverse restores the input."""
    from schroedinger_b200 import device as dev
    rng = np.random.default_rng(5)
    layout = dev.FrameLayout.yuv420("s16", 256, 128)
    count = 5
    a = dev.PictureSlab(layout, count)
    b = dev.PictureSlab(layout, count)
    c = dev.PictureSlab(layout, count)
    planes = {}
    for p in range(count):
        for k, (w, h) in enumerate(layout.comp_sizes):
            planes[p, k] = rng.integers(-255, 256, size=(h, w)).astype(np.int16)
            a.upload(p, k, planes[p, k])
    dev.iwt_forward(a, b, 0, 3)
    dev.iwt_inverse(b, c, 0, 3)
    for p in range(count):
        for k in range(3):
            want = helpers.cpu_wavelet(ORACLE, "oracle", "fwd", planes[p, k].copy(), 0, 3)
            assert np.array_equal(b.download(p, k), want)
            assert np.array_equal(c.download(p, k), planes[p, k])


@pytest.mark.parametrize("dtype", [np.int16, np.int32])
@pytest.mark.parametrize("filt", range(7))
def test_fast_path_shapes(cuda, filt, dtype):
    """Sizes whose half-width / half-height are multiples of 16 take the register-chunk
    kernels: cover single-chunk planes (both edges in one chunk), partial tiles, several
    tiles in both directions, and full-range values that wrap."""
    rng = np.random.default_rng(300 + filt)
    amp_full = 32767 if dtype == np.int16 else 2 ** 31 - 1
    for (h, w) in ((32, 32), (32, 64), (64, 32), (96, 160), (128, 288), (192, 416), (544, 960)):
        for amp in (255, amp_full):
            a = rng.integers(-amp, amp + 1, size=(h, w)).astype(dtype)
            for d in ("inv", "fwd"):
                want = helpers.cpu_wavelet(ORACLE, "oracle", d, a.copy(), filt)
                got = gpu_iwt(d, [a], filt, 1)[0]
                assert np.array_equal(got, want), (filt, dtype, (h, w), amp, d)


@pytest.mark.parametrize("dtype", [np.int16, np.int32])
@pytest.mark.parametrize("filt", range(7))
def test_generic_kernel_forced_on_fast_path_shapes(cuda, filt, dtype):
    """sb2_iwt_force_generic: the generic tile kernel on sizes the register-chunk kernels normally take
    (so both implementations are checked on the same input), and the launch tags say which one ran."""
    from schroedinger_b200 import lib
    from tests.test_hbm_gpu import launched_tags
    rng = np.random.default_rng(900 + filt)
    for (h, w) in ((32, 64), (128, 288), (544, 960)):
        a = rng.integers(-2000, 2001, size=(h, w)).astype(dtype)
        for d in ("inv", "fwd"):
            want = helpers.cpu_wavelet(ORACLE, "oracle", d, a.copy(), filt)
            got = {}
            tags = {}
            for forced in (0, 1):
                lib.sb2_iwt_force_generic(forced)
                try:
                    tags[forced] = launched_tags(lambda: got.__setitem__(forced, gpu_iwt(d, [a], filt, 1)[0]))
                finally:
                    lib.sb2_iwt_force_generic(0)
                assert np.array_equal(got[forced], want), (filt, dtype, (h, w), d, forced)
            assert all(t.endswith("_generic") for t in tags[1]) and tags[1]
            assert not any(t.endswith("_generic") for t in tags[0]) and tags[0]


@pytest.mark.parametrize("dtype", [np.int16, np.int32])
@pytest.mark.parametrize("filt", [0, 2, 6])
def test_fused_last_two_levels(cuda, filt, dtype):
    """Levels 1 + 0 of the inverse as one fused launch (wavelet_inv_fused2_kernel) against the oracle and
    against the one-launch-per-level path on the same input: single-tile planes (both picture edges in
    one window), partial tiles, many tiles, 4:2:0 frames, depths 2..4, in place, full-range values."""
    from schroedinger_b200 import lib
    from tests.test_hbm_gpu import launched_tags
    rng = np.random.default_rng(700 + filt)
    amp_full = 32767 if dtype == np.int16 else 2 ** 31 - 1
    npx = 40 if dtype == np.int16 else 36                     # smallest plane: n/2 >= NPX, m/2 >= 20
    w0 = (4 * npx + 31) // 32 * 32
    shapes = [((96, w0), 2), ((128, w0 + 32), 2), ((160, 352), 3), ((288, 544), 2), ((544, 960), 4), ((1088, 1920), 3)]
    for (h, w), depth in shapes:
        for amp in (600, amp_full):
            planes = [rng.integers(-amp, amp + 1, size=(h, w)).astype(dtype)]
            if (h // 2) % (1 << depth) == 0 and (w // 2) % (1 << depth) == 0 and h >= 288:
                planes += [rng.integers(-amp, amp + 1, size=(h // 2, w // 2)).astype(dtype) for _ in range(2)]
            want = [helpers.cpu_wavelet(ORACLE, "oracle", "inv", p.copy(), filt, depth) for p in planes]
            for fused in (1, 0):
                lib.sb2_iwt_enable_fused(fused)
                try:
                    got = {}
                    tags = launched_tags(lambda: got.__setitem__(0, gpu_iwt("inv", planes, filt, depth)))
                    got_ip = gpu_iwt("inv", planes, filt, depth, in_place=True)
                finally:
                    lib.sb2_iwt_enable_fused(0)
                for c in range(len(planes)):
                    assert np.array_equal(got[0][c], want[c]), (filt, dtype, (h, w), depth, amp, fused, c)
                    assert np.array_equal(got_ip[c], want[c]), (filt, dtype, (h, w), depth, amp, fused, c, "in place")
                # the fused kernel needs every component to be eligible (chroma of the small frames is not)
                eligible = all(p.shape[1] // 4 >= npx and p.shape[0] // 4 >= 20 and p.shape[1] % 32 == 0 and p.shape[0] % 32 == 0
                               for p in planes)
                assert any(t.endswith("_fused2") for t in tags) == bool(fused and eligible), (tags, fused, eligible, (h, w))
