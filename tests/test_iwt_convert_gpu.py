"""GPU parity: the inverse transform with the decoder's combine step fused into its last level
(sb2_iwt_inverse_convert: shift right + convert to 8 bits + crop in the level-0 kernel's epilogue) against the
oracle's inverse transform followed by the reference's shift and convert semantics (numpy shift, oracle convert)."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()


def np_shift_right(a, shift):
    if not shift:
        return a
    rnd = (1 << shift) >> 1
    return ((a.astype(np.int64) + rnd).astype(a.dtype).astype(np.int64) >> shift).astype(a.dtype)


def want_pictures(planes, filt, depth, shift, pw, ph):
    sd = 2 if planes[0].dtype == np.int32 else 1
    inv = [np_shift_right(helpers.cpu_wavelet(ORACLE, "oracle", "inv", p.copy(), filt, depth), shift) for p in planes]
    h, w = planes[0].shape
    return helpers.oracle_convert(ORACLE, inv, sd, w, h, 0, pw, ph)


@pytest.mark.parametrize("dtype", [np.int16, np.int32])
@pytest.mark.parametrize("filt", range(7))
def test_inverse_convert_matches_oracle(cuda, filt, dtype):
    from schroedinger_b200 import device as dev
    rng = np.random.default_rng(500 + filt)
    name = "s32" if dtype == np.int32 else "s16"
    info = np.iinfo(dtype)
    # (iwt width, height, depth, picture width, height, shift, amplitude): crops in both directions, the 1080p shape,
    # values that exercise the saturation and -- full range -- the wrap-around of the shift and the converters
    cases = [(64, 32, 1, 64, 32, 0, 300), (192, 96, 2, 180, 90, 0, 600), (192, 96, 2, 192, 96, 2, 2000),
             (352, 288, 3, 352, 288, 0, info.max), (1920, 1088, 4, 1920, 1080, 0, 400), (1920, 1088, 4, 1920, 1080, 2, 1500)]
    for (w, h, depth, pw, ph, shift, amp) in cases:
        if (w // 2) % (1 << depth) or (h // 2) % (1 << depth):
            continue
        planes = [rng.integers(-amp, amp + 1, size=s, dtype=np.int64).astype(dtype) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2))]
        src = dev.PictureSlab(dev.FrameLayout.yuv420(name, w, h), 2)
        out = dev.PictureSlab(dev.FrameLayout.yuv420("u8", pw, ph), 2)
        out.buf.fill_(0x5a)
        for p in range(2):
            for c in range(3):
                src.upload(p, c, planes[c] if p == 0 else planes[c][::-1].copy())
        dev.iwt_inverse_convert(src, out, filt, depth, shift)
        want = want_pictures(planes, filt, depth, shift, pw, ph)
        want2 = want_pictures([a[::-1].copy() for a in planes], filt, depth, shift, pw, ph)
        for c in range(3):
            assert np.array_equal(out.download(0, c), want[c]), (filt, dtype, (w, h), shift, c)
            assert np.array_equal(out.download(1, c), want2[c]), (filt, dtype, (w, h), shift, c, "second picture")


def test_inverse_convert_leaves_the_border_and_padding_alone(cuda):
    """A picture slab with an edge extension: only the picture's pixels are written."""
    from schroedinger_b200 import device as dev
    rng = np.random.default_rng(9)
    w, h, depth = 192, 96, 2
    planes = [rng.integers(-500, 501, size=s).astype(np.int16) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2))]
    src = dev.PictureSlab(dev.FrameLayout.yuv420("s16", w, h), 1)
    out = dev.PictureSlab(dev.FrameLayout.yuv420("u8", 180, 90, 32, False), 1)
    out.buf.fill_(0x77)
    for c in range(3):
        src.upload(0, c, planes[c])
    dev.iwt_inverse_convert(src, out, 0, depth)
    want = want_pictures(planes, 0, depth, 0, 180, 90)
    for c in range(3):
        full = out.download(0, c, with_border=True)
        assert np.array_equal(full[32:-32, 32:-32], want[c]), c
        full[32:-32, 32:-32] = 0x77
        assert (full == 0x77).all(), c


def test_inverse_convert_refuses_what_the_fast_kernels_do_not_cover(cuda):
    from schroedinger_b200 import device as dev
    from schroedinger_b200._lib import Sb2Error
    src = dev.PictureSlab(dev.FrameLayout("s16", [(40, 24)]), 1)          # half sizes not multiples of 8
    out = dev.PictureSlab(dev.FrameLayout("u8", [(40, 24)]), 1)
    with pytest.raises(Sb2Error):
        dev.iwt_inverse_convert(src, out, 0, 1)


@pytest.mark.parametrize("domain_kind", ["malloc", "cuda"])
@pytest.mark.parametrize("shape", [(480, 288, 4, 480, 270), (104, 72, 2, 100, 70)])       # the second is not covered by the fused kernel
def test_inverse_iwt_combine_drop_in(cuda, domain_kind, shape):
    """schro_b200_frame_inverse_iwt_combine against schro_frame_inverse_iwt_transform + schro_frame_shift_right +
    schro_frame_convert semantics (oracle), on host and CUDA-domain frames, fused and fall-back shapes."""
    import ctypes
    from schroedinger_b200 import compat, lib
    w, h, depth, pw, ph = shape
    rng = np.random.default_rng(w)
    planes = [rng.integers(-900, 901, size=s).astype(np.int16) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2))]
    params = compat.make_params(pw, ph, wavelet_filter_index=2, transform_depth=depth, iwt_luma_width=w, iwt_luma_height=h)
    params.iwt_chroma_width, params.iwt_chroma_height = w // 2, h // 2
    for shift in (0, 1):
        want = want_pictures(planes, 2, depth, shift, pw, ph)
        hf = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, w, h, 0, 0)
        for c in range(3):
            compat.frame_plane(hf, c)[...] = planes[c]
        ho = compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, pw, ph, 0, 0)
        if domain_kind == "cuda":
            dom = compat.cuda_domain()
            f = compat.frame_new_and_alloc(dom, compat.FORMAT_S16_420, w, h, 0, 0)
            o = compat.frame_new_and_alloc(dom, compat.FORMAT_U8_420, pw, ph, 0, 0)
            lib.schro_frame_to_gpu(f, hf)
            lib.schro_b200_frame_inverse_iwt_combine(o, f, ctypes.byref(params), shift)
            lib.schro_gpuframe_to_cpu(ho, o)
            back = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, w, h, 0, 0)
            lib.schro_gpuframe_to_cpu(back, f)
            for c in range(3):
                assert np.array_equal(np.array(compat.frame_plane(back, c)), planes[c]), ("coefficients untouched", c)
            for fr in (f, o, back):
                lib.schro_frame_unref(fr)
        else:
            lib.schro_b200_frame_inverse_iwt_combine(ho, hf, ctypes.byref(params), shift)
            for c in range(3):
                assert np.array_equal(np.array(compat.frame_plane(hf, c)), planes[c]), ("coefficients untouched", c)
        for c in range(3):
            assert np.array_equal(np.array(compat.frame_plane(ho, c)), want[c]), (shape, domain_kind, shift, c)
        lib.schro_frame_unref(hf)
        lib.schro_frame_unref(ho)
