"""GPU parity: the rough (bigblock) motion search -- sb2_rough_scan_nohint / _hint and the
schro_rough_me_* drop-ins -- against the oracle and the compiled reference's golden fields, bit-exact."""
import ctypes
import os

import numpy as np
import pytest
import torch

from tests import helpers
from tests.golden import make_golden as mg

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()
GOLD = np.load(os.path.join(helpers.GOLDEN_DIR, "rough.npz"))


@pytest.fixture(params=["staged", "unstaged"], autouse=True)
def full_search_path(request):
    """Every test runs twice: the full search with its windows staged in shared memory (default) and
    with every row segment read from global memory (sb2_rough_force_unstaged)."""
    from schroedinger_b200 import lib
    lib.sb2_rough_force_unstaged(1 if request.param == "unstaged" else 0)
    yield request.param
    lib.sb2_rough_force_unstaged(0)


def gpu_rough(pairs, width, height, levels, ref_index=0, xbsep=8, ybsep=8, dists=(12, 4)):
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
    prm = dev.HbmParams(xbsep, ybsep, nbx, nby, ref_index, 0, 1, 1)
    fields = dev.rough_scan(prm, ps, pr, *dists)
    torch.cuda.synchronize()
    out = []
    for p in range(count):
        f = [np.zeros(nbx * nby, helpers.MV_DTYPE)]
        f += [fl.cpu().numpy().view(helpers.MV_DTYPE).reshape(count, nbx * nby)[p] for fl in fields[1:]]
        out.append(np.stack(f))
    return out


def check(got, want, levels, what):
    for l in range(levels, 0, -1):
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(got[l][f], want[l][f]), (what, l, f)


def test_rough_golden(cuda):
    for idx, (w, h, lv, ri, pan, dn, dh) in enumerate(mg.ROUGH_GOLDEN_CASES):
        s, r = helpers.panning_pair(w, h, np.random.default_rng(3000 + idx), pan)
        got = gpu_rough([(s, r)], w, h, lv, ri, dists=(dn, dh))
        check(got[0], GOLD[f"r{idx}_fields"], lv, idx)


@pytest.mark.parametrize("w,h,levels,bs,dists", [
    (176, 144, 2, (8, 8), (12, 4)),        # ragged: blocks outside the level's frame, partial blocks
    (360, 270, 3, (8, 8), (12, 4)),
    (200, 72, 3, (8, 8), (12, 4)),
    (100, 70, 2, (8, 8), (12, 4)),         # rows that are not 8-byte aligned
    (384, 288, 2, (12, 12), (12, 4)),      # other block sizes take the position-per-lane path
    (256, 256, 2, (16, 8), (9, 3)),
    (320, 192, 3, (8, 8), (20, 7)),        # the largest window SchroMetricScan allows; hint windows beyond the register path
    (320, 192, 2, (8, 8), (3, 1)),
])
def test_rough_vs_oracle(cuda, w, h, levels, bs, dists):
    rng = np.random.default_rng(w * 7 + h)
    s, r = helpers.panning_pair(w, h, rng, (6, -3))
    want, _, _ = helpers.oracle_rough(ORACLE, s, r, w, h, bs[0], bs[1], levels, 0, *dists)
    got = gpu_rough([(s, r)], w, h, levels, 0, bs[0], bs[1], dists)
    check(got[0], want, levels, (w, h))


def test_rough_batch_and_incoherent_content(cuda):
    """Several pairs in one launch, among them flat pictures (every SAD ties) and pure noise."""
    w, h, levels = 256, 128, 3
    rng = np.random.default_rng(77)
    flat = [np.full((h, w), 77, np.uint8), np.full((h // 2, w // 2), 10, np.uint8), np.full((h // 2, w // 2), 200, np.uint8)]
    noisy = [rng.integers(0, 256, size=a.shape).astype(np.uint8) for a in flat]
    pairs = [(flat, flat), (noisy, flat), (noisy, [a[::-1].copy() for a in noisy]),
             helpers.panning_pair(w, h, rng, (-9, 7)), helpers.panning_pair(w, h, rng, (1, 0))]
    got = gpu_rough(pairs, w, h, levels, ref_index=1)
    for p, (s, r) in enumerate(pairs):
        want, _, _ = helpers.oracle_rough(ORACLE, s, r, w, h, levels=levels, ref_index=1)
        check(got[p], want, levels, p)


def test_rough_full_search_at_full_resolution(cuda):
    """The nohint function on level 0 of a 1080p picture (the throughput configuration bench.py times):
    checked against the oracle on a sample of block rows."""
    from schroedinger_b200 import device as dev
    w, h = 1920, 1080
    rng = np.random.default_rng(5)
    s, r = helpers.panning_pair(w, h, rng, (7, -4))
    ps, pr = dev.Pyramid(w, h, 1, 0), dev.Pyramid(w, h, 1, 0)
    for c in range(3):
        ps.slabs[0].upload(0, c, s[c])
        pr.slabs[0].upload(0, c, r[c])
    ps.build()
    pr.build()
    nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
    out = torch.empty(nbx * nby * 20, dtype=torch.uint8, device="cuda")
    dev.rough_scan_nohint(dev.HbmParams(8, 8, nbx, nby, 0, 0, 1, 1), ps.slabs[0], pr.slabs[0], 0, 12, out)
    got = out.cpu().numpy().view(helpers.MV_DTYPE).reshape(nby, nbx)
    hs, hr = helpers.build_pyramid(ORACLE, "oracle", s, 0), helpers.build_pyramid(ORACLE, "oracle", r, 0)
    want = np.zeros(nbx * nby, helpers.MV_DTYPE)
    ORACLE.oracle_rough_scan_nohint.restype = None
    a, b = helpers.pyr_level_struct(hs[0]), helpers.pyr_level_struct(hr[0])
    ORACLE.oracle_rough_scan_nohint(ctypes.byref(a), ctypes.byref(b), 8, 8, nbx, nby, 0, 0, 12,
                                    want.ctypes.data_as(ctypes.c_void_p))
    want = want.reshape(nby, nbx)
    for f in ("flags", "metric", "v"):
        assert np.array_equal(got[f], want[f]), f
    inner = got["v"][8:-8, 8:-8].astype(int)
    assert np.mean((inner[..., 0] == -7) & (inner[..., 2] == 4)) > 0.9


def test_rough_me_drop_in(cuda):
    """schro_rough_me_new_from_frames / _heirarchical_scan / direct motion_fields[] access / _free on
    malloc'd and on page-locked host frames."""
    from schroedinger_b200 import compat, lib
    from tests.test_host_api_gpu import _new_u8_frame
    w, h, levels = 320, 192, 3
    s, r = helpers.panning_pair(w, h, np.random.default_rng(31), (4, 6))
    want, _, _ = helpers.oracle_rough(ORACLE, s, r, w, h, levels=levels, ref_index=1)
    params = compat.make_params(w, h, xbsep=8, ybsep=8, xblen=12, yblen=12)
    n = params.x_num_blocks * params.y_num_blocks
    for domain in (None, compat.pinned_domain()):
        def pyramid(planes):
            frames = [_new_u8_frame(compat, lib, w, h, 32, True, planes, domain)]
            lib.schro_frame_mc_edgeextend(frames[0])
            cw, ch = w, h
            for _ in range(levels):
                cw, ch = (cw + 1) // 2, (ch + 1) // 2
                f = compat.frame_new_and_alloc(domain, compat.FORMAT_U8_420, cw, ch, 8, 0)
                lib.schro_frame_downsample(f, frames[-1])
                lib.schro_frame_mc_edgeextend(f)
                frames.append(f)
            return frames
        fs, fr = pyramid(s), pyramid(r)
        arr = compat.FrameP * (levels + 1)
        rme = lib.schro_rough_me_new_from_frames(None, None, ctypes.byref(params), 1, levels, arr(*fs), arr(*fr))
        lib.schro_rough_me_heirarchical_scan(rme)
        for l in range(levels, 0, -1):
            mf = rme.contents.motion_fields[l]
            got = np.ctypeslib.as_array(ctypes.cast(mf.contents.motion_vectors, ctypes.POINTER(ctypes.c_uint8)),
                                        shape=(n * 20,)).view(helpers.MV_DTYPE)
            for f in ("flags", "metric", "chroma_metric", "v"):
                assert np.array_equal(got[f], want[l][f]), (l, f)
        assert not rme.contents.motion_fields[0]
        # the level functions on their own, other distances
        lib.schro_rough_me_heirarchical_scan_nohint(rme, levels, 7)
        lib.schro_rough_me_heirarchical_scan_hint(rme, levels - 1, 2)
        want2, _, _ = helpers.oracle_rough(ORACLE, s, r, w, h, levels=levels, ref_index=1, nohint_distance=7,
                                           hint_distance=2)
        for l in (levels, levels - 1):
            mf = rme.contents.motion_fields[l]
            got = np.ctypeslib.as_array(ctypes.cast(mf.contents.motion_vectors, ctypes.POINTER(ctypes.c_uint8)),
                                        shape=(n * 20,)).view(helpers.MV_DTYPE)
            assert np.array_equal(got["v"], want2[l]["v"]) and np.array_equal(got["metric"], want2[l]["metric"]), l
        lib.schro_rough_me_free(rme)
        for f in fs + fr:
            lib.schro_frame_unref(f)
