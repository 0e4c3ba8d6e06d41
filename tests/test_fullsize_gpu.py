"""GPU parity at BASELINE.json's full sizes (2160p 4:2:0, the bench workload's geometry): every
stage of the picture core against the oracle on whole pictures, plus the size-independent
properties the domain offers (forward -> inverse round trip, a picture matched against itself,
zero-motion prediction reproducing its reference).  The oracle (plain C) needs a second or two
per stage at this size, so these stay in the default GPU suite."""
import numpy as np
import pytest

from tests import helpers
from tests.test_frame_gpu import gpu_upsampled, oracle_upsampled
from tests.test_hbm_gpu import gpu_hbm
from tests.test_obmc_gpu import gpu_obmc

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()

W, H, IWT_H = 3840, 2160, 2176          # schro_video_format: 2160 rows pad to 2176 for 5 levels


@pytest.mark.parametrize("filt,depth_name,levels", [(6, "s32", 5), (0, "s32", 5), (1, "s32", 5), (2, "s32", 5),
                                                    (3, "s32", 5), (4, "s32", 5), (5, "s32", 5),
                                                    (0, "s16", 4), (1, "s16", 4), (6, "s16", 4)])
def test_wavelet_2160p_oracle_and_round_trip(cuda, filt, depth_name, levels):
    """Two padded 2160p 4:2:0 pictures in one slab: forward == oracle on every plane, and the
    inverse gives the input back."""
    from schroedinger_b200 import device as dev
    rng = np.random.default_rng(2160 + filt)
    dt = np.int32 if depth_name == "s32" else np.int16
    layout = dev.FrameLayout.yuv420(depth_name, W, IWT_H)
    count = 2
    a, b, c = (dev.PictureSlab(layout, count) for _ in range(3))
    planes = {}
    for p in range(count):
        for k, (w, h) in enumerate(layout.comp_sizes):
            planes[p, k] = rng.integers(-255, 256, size=(h, w)).astype(dt)
            a.upload(p, k, planes[p, k])
    dev.iwt_forward(a, b, filt, levels)
    dev.iwt_inverse(b, c, filt, levels)
    for k in range(3):
        want = helpers.cpu_wavelet(ORACLE, "oracle", "fwd", planes[1, k].copy(), filt, levels)
        assert np.array_equal(b.download(1, k), want), (filt, depth_name, k)
    for p in range(count):
        for k in range(3):
            assert np.array_equal(c.download(p, k), planes[p, k]), (filt, depth_name, p, k)


@pytest.mark.parametrize("filt,depth_name,levels", [(f, "s32", 5) for f in range(7)] + [(0, "s16", 4), (1, "s16", 4), (6, "s16", 4)])
def test_wavelet_2160p_inverse_of_random_coefficients(cuda, filt, depth_name, levels):
    """The inverse on coefficients no forward transform produced (random, odd and even, every subband
    as loud as the others): the lifting steps' rounding shifts see both parities everywhere, which a
    forward -> inverse round trip never exercises.  Every plane of a padded 2160p picture == oracle."""
    from schroedinger_b200 import device as dev
    rng = np.random.default_rng(4320 + 10 * filt + (depth_name == "s16"))
    dt = np.int32 if depth_name == "s32" else np.int16
    layout = dev.FrameLayout.yuv420(depth_name, W, IWT_H)
    a, b = dev.PictureSlab(layout, 1), dev.PictureSlab(layout, 1)
    planes = []
    for k, (w, h) in enumerate(layout.comp_sizes):
        planes.append((rng.integers(-600, 601, size=(h, w)) | (rng.integers(0, 2, size=(h, w)))).astype(dt))
        a.upload(0, k, planes[k])
    dev.iwt_inverse(a, b, filt, levels)
    for k in range(3):
        want = helpers.cpu_wavelet(ORACLE, "oracle", "inv", planes[k].copy(), filt, levels)
        assert np.array_equal(b.download(0, k), want), (filt, depth_name, k)


def test_upsample_2160p_matches_oracle(cuda):
    rng = np.random.default_rng(12)
    imgs = [helpers.smooth_image(H, W, rng), rng.integers(0, 256, size=(H // 2, W // 2)).astype(np.uint8),
            helpers.smooth_image(H // 2, W // 2, rng)]
    got = gpu_upsampled(imgs, 32)
    for c in range(3):
        want = oracle_upsampled(imgs[c], 32)
        for p in range(4):
            assert np.array_equal(got[c][p], want[p]), (c, p)
        assert np.array_equal(got[c][0][32:-32, 32:-32], imgs[c])       # phase 0 is the picture itself


def test_block_matching_2160p_oracle_and_self_match(cuda):
    """The bench geometry (8x8 blocks, 4 levels): a panning pair against the oracle, and a picture
    matched against itself in the same launch -- every vector zero, every metric zero."""
    rng = np.random.default_rng(13)
    s, r = helpers.panning_pair(W, H, rng, (5, 3))
    noise = [rng.integers(0, 256, size=p.shape).astype(np.uint8) for p in s]
    got, _ = gpu_hbm([(s, r), (noise, noise)], W, H, 4)
    want, _, _ = helpers.oracle_hbm(ORACLE, s, r, W, H, levels=4)
    for f in ("flags", "metric", "chroma_metric", "v"):
        assert np.array_equal(got[0][f], want[f]), f
    assert not got[1]["v"].any()
    assert not got[1]["metric"].any()


def test_obmc_2160p_matches_oracle(cuda):
    """BASELINE config 4's renderer at the bench size: two references, 12x12 / 8x8 blocks,
    quarter-pel vectors, s32 residual (the inverse wavelet's output type in the bench)."""
    case = helpers.ObmcCase(ORACLE, W, H, rng=np.random.default_rng(14), res_is_s32=True)
    want = helpers.oracle_obmc(ORACLE, case, 1)
    got = gpu_obmc(case, 1)[0]
    for k in range(3):
        for q in range(3):
            assert np.array_equal(got[k][q], want[k][q]), (k, q)


def test_obmc_2160p_zero_motion_reproduces_reference(cuda):
    """One reference, zero vectors, zero residual: the OBMC weights of the blocks covering a pixel
    sum to 64, so the rendered picture is the reference picture, whatever the size."""
    case = helpers.ObmcCase(ORACLE, W, H, rng=np.random.default_rng(15), num_refs=1)
    case.mvs = np.zeros(case.nbx * case.nby, dtype=helpers.MV_DTYPE)
    case.mvs["flags"] = 1                       # every block predicted from reference 1, vector (0, 0)
    for k in range(3):
        case.residual[k][...] = 0
    got = gpu_obmc(case, 1)[0]
    for k in range(3):
        assert np.array_equal(got[k][2], case.ref0[k].phase(0, with_border=False)), k
