"""GPU: combine / convert glue (sb2_frame_convert, sb2_frame_add; schro_frame_convert / _add /
_subtract drop-ins) against the oracle, bit-exact."""
import ctypes

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()
NAMES = {0: "u8", 1: "s16", 2: "s32"}
SIZES = [((64, 48), (64, 48)), ((70, 50), (64, 48)), ((64, 48), (72, 56)), ((33, 17), (40, 24)), ((2, 2), (2, 2)),
         ((1920, 1080), (1920, 1088))]


def _slab(dev, depth, w, h, count, ext=0):
    return dev.PictureSlab(dev.FrameLayout.yuv420(NAMES[depth], w, h, ext), count)


@pytest.mark.parametrize("sdepth,ddepth", [(0, 1), (0, 2), (1, 0), (2, 0), (1, 2), (2, 1), (0, 0), (1, 1), (2, 2)])
def test_convert_matches_oracle(cuda, sdepth, ddepth):
    from schroedinger_b200 import device as dev
    rng = np.random.default_rng(100 + 10 * sdepth + ddepth)
    for (sw, sh), (dw, dh) in SIZES:
        count = 2
        s, d = _slab(dev, sdepth, sw, sh, count, ext=4), _slab(dev, ddepth, dw, dh, count)
        srcs = [helpers.random_planes(rng, sdepth, sw, sh, p == 0) for p in range(count)]
        for p in range(count):
            for c in range(3):
                s.upload(p, c, srcs[p][c])
        dev.frame_convert(s, d)
        for p in range(count):
            want = helpers.oracle_convert(ORACLE, srcs[p], sdepth, sw, sh, ddepth, dw, dh)
            for c in range(3):
                assert np.array_equal(d.download(p, c), want[c]), (sdepth, ddepth, (sw, sh), (dw, dh), p, c)


@pytest.mark.parametrize("sdepth", [0, 1])
@pytest.mark.parametrize("subtract", [False, True])
def test_add_subtract_match_oracle(cuda, sdepth, subtract):
    from schroedinger_b200 import device as dev
    rng = np.random.default_rng(7 + sdepth)
    for (sw, sh), (dw, dh) in SIZES:
        s, d = _slab(dev, sdepth, sw, sh, 1), _slab(dev, 1, dw, dh, 1, ext=2)
        src = helpers.random_planes(rng, sdepth, sw, sh, True)
        dst = helpers.random_planes(rng, 1, dw, dh, True)
        for c in range(3):
            s.upload(0, c, src[c])
            d.upload(0, c, dst[c])
        dev.frame_add(d, s, subtract)
        want = helpers.cpu_add(ORACLE, "oracle", dst, dw, dh, src, sdepth, sw, sh, subtract)
        for c in range(3):
            assert np.array_equal(d.download(0, c), want[c]), (sdepth, subtract, (sw, sh), (dw, dh), c)


@pytest.mark.parametrize("domain_kind", ["malloc", "cuda"])
def test_frame_convert_add_drop_in(cuda, domain_kind):
    """The decoder's tail for an intra picture: s16 wavelet output -> u8 picture
    (schro_frame_convert, schrodecoder.c), and the encoder's residual: s16 -= u8 prediction."""
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(42)
    w, h = 352, 288
    fmt = {0: compat.FORMAT_U8_420, 1: compat.FORMAT_S16_420}
    dom = compat.cuda_domain() if domain_kind == "cuda" else None
    src = helpers.random_planes(rng, 1, w, h + 8, False)            # padded s16 frame, cropped by the convert
    pred = helpers.random_planes(rng, 0, w, h, True)

    def frame(depth, fw, fh, planes=None):
        host = compat.frame_new_and_alloc(None, fmt[depth], fw, fh)
        if planes is not None:
            for c in range(3):
                compat.frame_plane(host, c)[...] = planes[c]
        if dom is None:
            return host, host
        f = compat.frame_new_and_alloc(dom, fmt[depth], fw, fh)
        lib.schro_frame_to_gpu(f, host)
        return f, host

    fs, _ = frame(1, w, h + 8, src)
    fd, hd = frame(0, w, h)
    lib.schro_frame_convert(fd, fs)
    if fd is not hd:
        lib.schro_gpuframe_to_cpu(hd, fd)
    want = helpers.oracle_convert(ORACLE, src, 1, w, h + 8, 0, w, h)
    for c in range(3):
        assert np.array_equal(compat.frame_plane(hd, c), want[c]), c
    fp, _ = frame(0, w, h, pred)
    lib.schro_frame_subtract(fs, fp)
    lib.schro_frame_add(fs, fp)
    lib.schro_frame_subtract(fs, fp)
    hs = compat.frame_new_and_alloc(None, fmt[1], w, h + 8)
    if dom is None:
        hs = fs
    else:
        lib.schro_gpuframe_to_cpu(hs, fs)
    want = helpers.cpu_add(ORACLE, "oracle", src, w, h + 8, pred, 0, w, h, True)
    for c in range(3):
        assert np.array_equal(compat.frame_plane(hs, c), want[c]), c
