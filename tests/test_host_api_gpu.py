"""GPU: the drop-in C API (schro_* symbols, host layer) with HOST buffers, against the oracle.
Reads like the reference's own testsuite/wavelet_2d.c: build a SchroFrameData, call
schro_wavelet_transform_2d / schro_wavelet_inverse_transform_2d, compare."""
import ctypes

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()


@pytest.mark.parametrize("dtype", [np.int16, np.int32])
@pytest.mark.parametrize("filt", range(7))
def test_wavelet_2d_host_buffers(cuda, filt, dtype):
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(filt)
    for (h, w) in ((20, 20), (2, 2), (40, 38), (64, 200)):
        a = rng.integers(-255, 256, size=(h, w)).astype(dtype)
        want_f = helpers.cpu_wavelet(ORACLE, "oracle", "fwd", a.copy(), filt)
        want_i = helpers.cpu_wavelet(ORACLE, "oracle", "inv", a.copy(), filt)
        b = a.copy()
        fd = compat.frame_data(b)
        lib.schro_wavelet_transform_2d(ctypes.byref(fd), filt, None)
        assert np.array_equal(b, want_f)
        c = a.copy()
        fd = compat.frame_data(c)
        lib.schro_wavelet_inverse_transform_2d(ctypes.byref(fd), ctypes.byref(fd), filt, None)
        assert np.array_equal(c, want_i)


@pytest.mark.parametrize("domain_kind", ["malloc", "pinned", "cuda"])
def test_frame_iwt_round_trip_all_domains(cuda, domain_kind):
    """schro_frame_iwt_transform / schro_frame_inverse_iwt_transform on frames from every
    memory domain (the reference's schro_frame_iwt_transform, schroframe.c:1192)."""
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(3)
    w, h, depth, filt = 352, 288, 4, 0
    params = compat.make_params(w, h, filt, depth)
    iw, ih = params.iwt_luma_width, params.iwt_luma_height
    host = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, iw, ih)
    planes = []
    for c in range(3):
        v = compat.frame_plane(host, c)
        v[...] = rng.integers(-255, 256, size=v.shape)
        planes.append(v.copy())
    if domain_kind == "malloc":
        work, dom = host, None
    else:
        dom = compat.pinned_domain() if domain_kind == "pinned" else compat.cuda_domain()
        work = compat.frame_new_and_alloc(dom, compat.FORMAT_S16_420, iw, ih)
        lib.schro_frame_to_gpu(work, host)
    lib.schro_frame_iwt_transform(work, ctypes.byref(params))
    if work is not host:
        lib.schro_gpuframe_to_cpu(host, work)
    for c in range(3):
        want = helpers.cpu_wavelet(ORACLE, "oracle", "fwd", planes[c].copy(), filt, depth)
        assert np.array_equal(compat.frame_plane(host, c), want), (domain_kind, c)
    lib.schro_frame_inverse_iwt_transform(work, ctypes.byref(params))
    if work is not host:
        lib.schro_gpuframe_to_cpu(host, work)
    for c in range(3):
        assert np.array_equal(compat.frame_plane(host, c), planes[c]), (domain_kind, c, "round trip")
    if work is not host:
        lib.schro_frame_unref(work)
        lib.schro_memory_domain_free(dom)
    lib.schro_frame_unref(host)


# ---------------------------------------------------------------------------------------
# frame preparation, OBMC and block matching through the drop-in C API, host frames
# ---------------------------------------------------------------------------------------
def _new_u8_frame(compat, lib, w, h, ext, upsampled, images, domain=None):
    f = compat.frame_new_and_alloc(domain, compat.FORMAT_U8_420, w, h, ext, 1 if upsampled else 0)
    for c in range(3):
        compat.frame_plane(f, c, with_border=True)[...] = 0x33
        compat.frame_plane(f, c)[...] = images[c]
    return f


def test_edgeextend_and_upsample_host_frame(cuda):
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(21)
    w, h = 176, 144
    imgs = [rng.integers(0, 256, size=s).astype(np.uint8) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2))]
    f = _new_u8_frame(compat, lib, w, h, 32, True, imgs)
    lib.schro_frame_mc_edgeextend(f)
    lib.schro_upsampled_frame_upsample(f)
    assert f.contents.upsample_done == 1
    for c in range(3):
        pl = helpers.HostPlane(imgs[c].shape[1], imgs[c].shape[0], ext=32, upsampled=True, fill=0x33)
        pl.set_image(imgs[c])
        helpers.cpu_edgeextend(ORACLE, "oracle", pl)
        helpers.cpu_upsample(ORACLE, "oracle", pl)
        for p in range(4):
            assert np.array_equal(compat.frame_plane(f, c, phase=p, with_border=True), pl.phase(p)), (c, p)
    lib.schro_frame_unref(f)


def test_upsample_horiz_vert_standalone(cuda):
    """schro_frame_upsample_horiz / _vert on bare planes, incl. the n <= 8 no-copy rule."""
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(5)
    for (h, w) in ((12, 20), (5, 8), (9, 7), (3, 30)):
        src = rng.integers(0, 256, size=(h, w)).astype(np.uint8)
        taps = np.array([-1, 3, -7, 21, 21, -7, 3, -1])
        wanth = np.zeros_like(src)
        wantv = np.zeros_like(src)
        for y in range(h):
            for x in range(w):
                acc = sum(int(taps[j]) * int(src[y, min(max(x + j - 3, 0), w - 1)]) for j in range(8))
                wanth[y, x] = min(max((acc + 16) >> 5, 0), 255)
                acc = sum(int(taps[j]) * int(src[min(max(y + j - 3, 0), h - 1), x]) for j in range(8))
                wantv[y, x] = min(max((acc + 16) >> 5, 0), 255)
        if w > 8:
            wanth[:, w - 1] = src[:, w - 1]
        wantv[h - 1, :] = src[h - 1, :]
        for fn, want in ((lib.schro_frame_upsample_horiz, wanth), (lib.schro_frame_upsample_vert, wantv)):
            dst = np.zeros_like(src)
            fs, fdst = compat.frame_data(src), compat.frame_data(dst)
            fn(ctypes.byref(fdst), ctypes.byref(fs))
            assert np.array_equal(dst, want), (h, w, fn.__name__)


def test_frame_downsample_host(cuda):
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(8)
    w, h = 100, 70
    imgs = [rng.integers(0, 256, size=s).astype(np.uint8) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2))]
    src = _new_u8_frame(compat, lib, w, h, 0, False, imgs)
    dst = compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, (w + 1) // 2, (h + 1) // 2, 8, 0)
    lib.schro_frame_downsample(dst, src)
    for c in range(3):
        assert np.array_equal(compat.frame_plane(dst, c), helpers.cpu_downsample(ORACLE, "oracle", imgs[c]))
    lib.schro_frame_unref(src)
    lib.schro_frame_unref(dst)


@pytest.mark.parametrize("domain_kind", ["malloc", "cuda"])
@pytest.mark.parametrize("add", [1, 0])
def test_motion_render_drop_in(cuda, add, domain_kind):
    """schro_motion_new + schro_motion_render exactly as the decoder / encoder call them
    (schrodecoder.c:1697-1792, schroencoder.c:2429-2460)."""
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(31)
    case = helpers.ObmcCase(ORACLE, 176, 144, rng=rng)
    want = helpers.oracle_obmc(ORACLE, case, add)
    params = compat.make_params(case.width, case.height, num_refs=2, xblen=case.xblen, yblen=case.yblen,
                                xbsep=case.xbsep, ybsep=case.ybsep, mv_precision=case.prec)
    dom = compat.cuda_domain() if domain_kind == "cuda" else None

    def to_domain(host):
        if dom is None:
            return host
        f = host.contents
        d = compat.frame_new_and_alloc(dom, f.format, f.width, f.height, f.extension, f.is_upsampled)
        lib.schro_frame_to_gpu(d, host)
        return d

    refs = []
    for planes in (case.ref0, case.ref1):
        f = _new_u8_frame(compat, lib, case.width, case.height, 32, True,
                          [p.phase(0, with_border=False) for p in planes])
        lib.schro_frame_mc_edgeextend(f)
        lib.schro_upsampled_frame_upsample(f)
        refs.append(to_domain(f))
    dest_h = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, case.width, case.height)
    addf_h = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, case.width, case.height)
    outf_h = compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, case.width, case.height, 32, 1)
    for c in range(3):
        compat.frame_plane(addf_h, c)[...] = case.residual[c]
        compat.frame_plane(dest_h, c)[...] = 0
    dest, addf, outf = to_domain(dest_h), to_domain(addf_h), to_domain(outf_h)
    motion = lib.schro_motion_new(ctypes.byref(params), refs[0], refs[1])
    ctypes.memmove(motion.contents.motion_vectors, case.mvs.ctypes.data, case.mvs.nbytes)
    lib.schro_motion_render(motion, dest, addf, add, outf if add else None)
    if dom is not None:
        lib.schro_gpuframe_to_cpu(dest_h, dest)
        lib.schro_gpuframe_to_cpu(addf_h, addf)
        lib.schro_gpuframe_to_cpu(outf_h, outf)
    for c in range(3):
        assert np.array_equal(compat.frame_plane(dest_h, c), want[c][0]), ("acc", c)
        if add:
            assert np.array_equal(compat.frame_plane(outf_h, c), want[c][2]), ("out", c)
        else:
            assert np.array_equal(compat.frame_plane(addf_h, c), want[c][1]), ("residual", c)
    lib.schro_motion_free(motion)


@pytest.mark.parametrize("add", [1, 0])
def test_motion_render_padded_addframe_1080p(cuda, add):
    """Both reference call sites pass a picture-size dest and an iwt-PADDED addframe (1080 -> 1088 rows
    at depth 4: schrodecoder.c:1784, schroencoder.c:2447); the rendered area is dest's
    (schromotion8.c:722-751).  Nothing may be written below row 1080: not into the addframe's padding,
    not into the output frame's border."""
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(32)
    w, h = 1920, 1080
    case = helpers.ObmcCase(ORACLE, w, h, rng=rng)
    want = helpers.oracle_obmc(ORACLE, case, add)
    params = compat.make_params(w, h, 0, 4, num_refs=2, xblen=case.xblen, yblen=case.yblen,
                                xbsep=case.xbsep, ybsep=case.ybsep, mv_precision=case.prec)
    iw, ih = params.iwt_luma_width, params.iwt_luma_height
    assert (iw, ih) == (1920, 1088)
    dom = compat.cuda_domain()

    def to_dev(host):
        f = host.contents
        d = compat.frame_new_and_alloc(dom, f.format, f.width, f.height, f.extension, f.is_upsampled)
        lib.schro_frame_to_gpu(d, host)
        return d

    refs = []
    for planes in (case.ref0, case.ref1):
        f = _new_u8_frame(compat, lib, w, h, 32, True, [p.phase(0, with_border=False) for p in planes])
        lib.schro_frame_mc_edgeextend(f)
        lib.schro_upsampled_frame_upsample(f)
        refs.append(to_dev(f))
    dest_h = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, w, h)
    addf_h = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, iw, ih)
    outf_h = compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, w, h, 32, 1)
    pad = []
    for c in range(3):
        a = compat.frame_plane(addf_h, c)
        a[...] = rng.integers(-3000, 3000, size=a.shape)            # padding rows carry a pattern
        a[:case.residual[c].shape[0], :] = case.residual[c]
        pad.append(a[case.residual[c].shape[0]:, :].copy())
        compat.frame_plane(dest_h, c)[...] = 0
    out_all = np.ctypeslib.as_array(ctypes.cast(outf_h.contents.regions[0], ctypes.POINTER(ctypes.c_uint8)),
                                    shape=(sum(outf_h.contents.components[k].length for k in range(3)),))
    out_all[...] = 0x5a
    out_before = out_all.copy()
    dest, addf, outf = to_dev(dest_h), to_dev(addf_h), to_dev(outf_h)
    motion = lib.schro_motion_new(ctypes.byref(params), refs[0], refs[1])
    ctypes.memmove(motion.contents.motion_vectors, case.mvs.ctypes.data, case.mvs.nbytes)
    lib.schro_motion_render(motion, dest, addf, add, outf if add else None)
    lib.schro_gpuframe_to_cpu(dest_h, dest)
    lib.schro_gpuframe_to_cpu(addf_h, addf)
    lib.schro_gpuframe_to_cpu(outf_h, outf)
    for c in range(3):
        rows = case.residual[c].shape[0]
        assert np.array_equal(compat.frame_plane(dest_h, c), want[c][0]), ("acc", c)
        assert np.array_equal(compat.frame_plane(addf_h, c)[rows:, :], pad[c]), ("addframe padding", c)
        if add:
            assert np.array_equal(compat.frame_plane(outf_h, c), want[c][2]), ("out", c)
        else:
            assert np.array_equal(compat.frame_plane(addf_h, c)[:rows, :], want[c][1]), ("residual", c)
    if add:
        # everything outside the three picture areas of the output frame is as it was
        touched = out_all != out_before
        inside = np.zeros_like(touched)
        base = outf_h.contents.regions[0]
        for k in range(3):
            comp = outf_h.contents.components[k]
            off = comp.data - base
            for y in range(comp.height):
                inside[off + y * comp.stride: off + y * comp.stride + comp.width] = True
        assert not (touched & ~inside).any()
    lib.schro_motion_free(motion)
    for f in refs + [dest, addf, outf, dest_h, addf_h, outf_h]:
        lib.schro_frame_unref(f)
    lib.schro_memory_domain_free(dom)


def test_hbm_drop_in(cuda):
    """Pyramid with schro_frame_downsample + schro_frame_mc_edgeextend (schroanalysis.c:9-28),
    then schro_hbm_scan and the level-0 refinement (schromotionest.c:76-77, 123-127)."""
    from schroedinger_b200 import compat, lib
    w, h, levels = 320, 192, 3
    s, r = helpers.panning_pair(w, h, np.random.default_rng(12), (4, -2))
    want, _, _ = helpers.oracle_hbm(ORACLE, s, r, w, h, levels=levels)
    params = compat.make_params(w, h, xbsep=8, ybsep=8, xblen=12, yblen=12)

    def pyramid(planes):
        frames = [_new_u8_frame(compat, lib, w, h, 32, True, planes)]
        lib.schro_frame_mc_edgeextend(frames[0])
        cw, ch = w, h
        for _ in range(levels):
            cw, ch = (cw + 1) // 2, (ch + 1) // 2
            f = compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, cw, ch, 8, 0)
            lib.schro_frame_downsample(f, frames[-1])
            lib.schro_frame_mc_edgeextend(f)
            frames.append(f)
        return frames

    fs, fr = pyramid(s), pyramid(r)
    arr = compat.FrameP * (levels + 1)
    hbm = lib.schro_hbm_new_from_frames(ctypes.byref(params), 0, levels, 0, arr(*fs), arr(*fr))
    lib.schro_hbm_scan(hbm)
    lib.schro_hierarchical_bm_scan_hint(hbm, 0, 3)
    n = params.x_num_blocks * params.y_num_blocks
    for l in range(levels + 1):
        mf = lib.schro_hbm_motion_field(hbm, l)
        got = np.ctypeslib.as_array(ctypes.cast(mf.contents.motion_vectors, ctypes.POINTER(ctypes.c_uint8)),
                                    shape=(n * 20,)).view(helpers.MV_DTYPE)
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(got[f], want[l][f]), (l, f)
    lib.schro_hbm_unref(hbm)


def test_metric_primitives_host(cuda):
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, size=(40, 64)).astype(np.uint8)
    b = rng.integers(0, 256, size=(40, 64)).astype(np.uint8)
    c = rng.integers(0, 256, size=(40, 64)).astype(np.uint8)
    for (w, h) in ((8, 8), (12, 12), (16, 5), (7, 3)):
        want = int(np.abs(a[:h, :w].astype(int) - b[:h, :w].astype(int)).sum())
        assert lib.schro_metric_absdiff_u8(a.ctypes.data, a.strides[0], b.ctypes.data, b.strides[0], w, h) == want
        fa, fb, fc = compat.frame_data(a), compat.frame_data(b), compat.frame_data(c)
        lib.schro_metric_get.restype = ctypes.c_int
        assert lib.schro_metric_get(ctypes.byref(fa), ctypes.byref(fb), w, h) == want
        lib.schro_metric_get_dc.restype = ctypes.c_int
        assert lib.schro_metric_get_dc(ctypes.byref(fa), 77, w, h) == int(np.abs(77 - a[:h, :w].astype(int)).sum())
        lib.schro_metric_get_biref.restype = ctypes.c_int
        x = (b[:h, :w].astype(int) * 3 + c[:h, :w].astype(int) * 1 + 2) >> 2
        assert lib.schro_metric_get_biref(ctypes.byref(fa), ctypes.byref(fb), 3, ctypes.byref(fc), 1, 2, w, h) \
            == int(np.abs(a[:h, :w].astype(int) - x).sum())
