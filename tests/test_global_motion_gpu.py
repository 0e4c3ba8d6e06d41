"""GPU parity: the reference's per-pixel renderer, which schro_motion_render uses when the picture has global
motion (sb2_obmc_render_ref, and the schro_motion_render / schro_motion_render_ref drop-ins), against the oracle
(pinned to the compiled reference in tests/test_oracle_global_motion.py), bit-exact, both directions."""
import ctypes

import numpy as np
import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()


def gpu_render_ref(case, gm, add, count=1):
    from schroedinger_b200 import device as dev
    sizes = case.comp_sizes
    ref_lay = dev.FrameLayout("u8", sizes, 32, True)
    refs = []
    for planes in (case.ref0, case.ref1):
        if planes is None:
            refs.append(None)
            continue
        slab = dev.PictureSlab(ref_lay, count)
        for p in range(count):
            for c, pl in enumerate(planes):
                start = p * ref_lay.pitch + ref_lay.offset[c] - pl.origin
                slab.buf[start:start + pl.buf.size].copy_(torch.from_numpy(pl.buf.reshape(-1)))
        refs.append(slab)
    res = dev.PictureSlab(dev.FrameLayout("s16", sizes), count)
    acc = dev.PictureSlab(dev.FrameLayout("s16", sizes), count)
    out = dev.PictureSlab(dev.FrameLayout("u8", sizes), count)
    for p in range(count):
        for c in range(3):
            res.upload(p, c, case.residual[c])
    mvs = torch.from_numpy(np.tile(case.mvs.view(np.uint8), count)).cuda()
    prm = dev.ObmcParams(case.xbsep, case.ybsep, case.xblen, case.yblen, case.nbx, case.nby,
                         case.prec, case.weights[0], case.weights[1], case.weights[2], case.hs, case.vs)
    dev.obmc_render_ref(prm, gm, mvs, refs[0], refs[1], res, add, out=out, acc=acc)
    torch.cuda.synchronize()
    return [[(acc.download(p, c), res.download(p, c), out.download(p, c)) for c in range(3)] for p in range(count)]


@pytest.mark.parametrize("prec", [0, 1, 2, 3])
@pytest.mark.parametrize("add", [1, 0])
def test_global_motion_vs_oracle(cuda, prec, add):
    rng = np.random.default_rng(60 + prec)
    case, gm = helpers.global_motion_case(ORACLE, 176, 144, rng, prec=prec)
    want = helpers.oracle_obmc_ref(ORACLE, case, add, gm)
    got = gpu_render_ref(case, gm, add, count=2)
    for p in range(2):
        for k in range(3):
            for part in range(3 if add else 2):
                assert np.array_equal(got[p][k][part], want[k][part]), (prec, add, p, k, part)


def test_global_motion_geometries_weights_outliers(cuda):
    rng = np.random.default_rng(8)
    for kw in (dict(xbsep=8, ybsep=8, xblen=8, yblen=8), dict(xbsep=16, ybsep=16, xblen=24, yblen=24, weights=(3, 1, 2)),
               dict(xbsep=4, ybsep=4, xblen=8, yblen=8, num_refs=1), dict(weights=(1, 3, 2), span=300, outliers=0.05)):
        case, gm = helpers.global_motion_case(ORACLE, 100, 70, rng, **kw)
        for add in (1, 0):
            want = helpers.oracle_obmc_ref(ORACLE, case, add, gm)
            got = gpu_render_ref(case, gm, add)[0]
            for k in range(3):
                for part in range(3 if add else 2):
                    assert np.array_equal(got[k][part], want[k][part]), (kw, add, k, part)


@pytest.mark.parametrize("add", [1, 0])
def test_schro_motion_render_with_global_motion(cuda, add):
    """The drop-in dispatcher: params->have_global_motion routes schro_motion_render to the per-pixel renderer."""
    from schroedinger_b200 import compat, lib
    from tests.test_host_api_gpu import _new_u8_frame
    rng = np.random.default_rng(33)
    case, gm = helpers.global_motion_case(ORACLE, 176, 144, rng)
    want = helpers.oracle_obmc_ref(ORACLE, case, add, gm)
    params = compat.make_params(case.width, case.height, num_refs=2, xblen=case.xblen, yblen=case.yblen,
                                xbsep=case.xbsep, ybsep=case.ybsep, mv_precision=case.prec)
    params.have_global_motion = 1
    names = ("b0", "b1", "a_exp", "a00", "a01", "a10", "a11", "c_exp", "c0", "c1")
    for r in range(2):
        for i, n in enumerate(names):
            setattr(params.global_motion[r], n, gm[10 * r + i])
    refs = []
    for planes in (case.ref0, case.ref1):
        f = _new_u8_frame(compat, lib, case.width, case.height, 32, True, [p.phase(0, with_border=False) for p in planes])
        lib.schro_frame_mc_edgeextend(f)
        lib.schro_upsampled_frame_upsample(f)
        refs.append(f)
    dest = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, case.width, case.height)
    addf = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, case.width, case.height)
    outf = compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, case.width, case.height, 32, 1)
    for c in range(3):
        compat.frame_plane(addf, c)[...] = case.residual[c]
        compat.frame_plane(dest, c)[...] = 0
    motion = lib.schro_motion_new(ctypes.byref(params), refs[0], refs[1])
    ctypes.memmove(motion.contents.motion_vectors, case.mvs.ctypes.data, case.mvs.nbytes)
    lib.schro_motion_render(motion, dest, addf, add, outf if add else None)
    for c in range(3):
        assert np.array_equal(compat.frame_plane(dest, c), want[c][0]), ("acc", c)
        if add:
            assert np.array_equal(compat.frame_plane(outf, c), want[c][2]), ("out", c)
        else:
            assert np.array_equal(compat.frame_plane(addf, c), want[c][1]), ("residual", c)
    lib.schro_motion_free(motion)
    for f in refs + [dest, addf, outf]:
        lib.schro_frame_unref(f)
