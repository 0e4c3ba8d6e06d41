"""GPU parity: OBMC motion-compensation renderer (sb2_obmc_render) against the oracle and
the golden outputs of the unmodified reference, bit-exact, both directions."""
import os

import numpy as np
import pytest
import torch

from tests import helpers
from tests.golden import make_golden as mg

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()
GOLD = np.load(os.path.join(helpers.GOLDEN_DIR, "motion.npz"))
KERNELS = {"auto": 0, "tma": 1, "scatter": 2, "pixel": 3, "blocks": 4}
RAN = {k: 0 for k in KERNELS}


@pytest.fixture(params=list(KERNELS), autouse=True)
def obmc_kernel(request):
    """Every test of this file runs four times: with the kernel the library picks and with each of the
    three kernels forced (block per warp on TMA-staged regions, block-major scatter, one thread per pixel).  A forced kernel
    that does not cover a case's geometry makes gpu_obmc return None and the case is skipped for it."""
    from schroedinger_b200 import lib
    lib.sb2_obmc_force_kernel(KERNELS[request.param])
    yield request.param
    lib.sb2_obmc_force_kernel(0)


def gpu_obmc(case, add, count=1):
    """Run `count` copies of the case as one slab; returns per picture, per component
    (acc, residual_after, out)."""
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
                assert pl.stride == ref_lay.stride[c]
                # whole plane incl. borders and all four phases, byte for byte
                start = p * ref_lay.pitch + ref_lay.offset[c] - pl.origin
                slab.buf[start:start + pl.buf.size].copy_(torch.from_numpy(pl.buf.reshape(-1)))
        refs.append(slab)
    rdepth = "s32" if case.res_is_s32 else "s16"
    res = dev.PictureSlab(dev.FrameLayout(rdepth, sizes), count)
    acc = dev.PictureSlab(dev.FrameLayout("s16", sizes), count)
    out = dev.PictureSlab(dev.FrameLayout("u8", sizes), count)
    for p in range(count):
        for c in range(3):
            res.upload(p, c, case.residual[c])
    mvs = torch.from_numpy(np.tile(case.mvs.view(np.uint8), count)).cuda()
    prm = dev.ObmcParams(case.xbsep, case.ybsep, case.xblen, case.yblen, case.nbx, case.nby,
                         case.prec, case.weights[0], case.weights[1], case.weights[2], case.hs, case.vs)
    from schroedinger_b200 import lib, Sb2Error
    try:
        dev.obmc_render(prm, mvs, refs[0], refs[1], res, add, out=out, acc=acc)
    except Sb2Error as ex:
        if "does not cover this geometry" in str(ex):
            return None
        raise
    which = lib.sb2_obmc_last_kernel()
    RAN[{1: "tma", 2: "scatter", 3: "pixel", 4: "blocks"}[which]] += 1
    return [[(acc.download(p, c), res.download(p, c), out.download(p, c)) for c in range(3)]
            for p in range(count)]


def test_obmc_golden(cuda):
    for idx, kw in enumerate(mg.OBMC_GOLDEN_CASES):
        for add in (1, 0):
            if not add and kw.get("res_is_s32"):
                continue
            case = helpers.ObmcCase(ORACLE, rng=np.random.default_rng(1000 + idx), **kw)
            got = gpu_obmc(case, add)
            if got is None:
                continue
            got = got[0]
            for k in range(3):
                for q, name in enumerate(("acc", "resid", "out")):
                    if q == 2 and not add:
                        continue
                    assert np.array_equal(got[k][q], GOLD[f"c{idx}_add{add}_k{k}_{name}"]), (kw, add, k, name)


@pytest.mark.parametrize("kw", [
    dict(width=100, height=60, xbsep=8, ybsep=8, xblen=8, yblen=8),
    dict(width=64, height=48, weights=(1, 2, 2)),
    dict(width=64, height=48, num_refs=1, weights=(2, 1, 1)),
    dict(width=80, height=48, chroma_format=1),
    dict(width=352, height=288, span=200, outliers=0.05),
    dict(width=64, height=48, xbsep=8, ybsep=4, xblen=12, yblen=8, prec=3),
    dict(width=40, height=24, xbsep=4, ybsep=4, xblen=8, yblen=8, prec=1, span=2000),
])
def test_obmc_matches_oracle(cuda, kw):
    for add in (1, 0):
        case = helpers.ObmcCase(ORACLE, rng=np.random.default_rng(77), **kw)
        want = helpers.oracle_obmc(ORACLE, case, add)
        got = gpu_obmc(case, add, count=2)
        if got is None:
            continue
        for p in range(2):
            for k in range(3):
                for q in range(3 if add else 2):
                    assert np.array_equal(got[p][k][q], want[k][q]), (kw, add, p, k, q)


def test_obmc_1080p_config4(cuda):
    """BASELINE config 4: 1080p, 12x12/8x8 blocks, quarter-pel, two references."""
    case = helpers.ObmcCase(ORACLE, 1920, 1080, rng=np.random.default_rng(4))
    for add in (1, 0):
        want = helpers.oracle_obmc(ORACLE, case, add)
        got = gpu_obmc(case, add)
        assert got is not None                      # the codec's default geometry runs on every kernel
        got = got[0]
        for k in range(3):
            for q in range(3 if add else 2):
                assert np.array_equal(got[k][q], want[k][q]), (add, k, q)


def test_obmc_every_kernel_ran(cuda, obmc_kernel):
    """(last in the file) the forced runs above really went through the kernel they name, and the
    default choice for the codec's geometries is the scatter kernel (the faster of the two)"""
    from schroedinger_b200 import lib
    case = helpers.ObmcCase(ORACLE, 96, 64, rng=np.random.default_rng(5))
    assert gpu_obmc(case, 1) is not None
    want = {"auto": 2, "tma": 1, "scatter": 2, "pixel": 3, "blocks": 4}[obmc_kernel]
    assert lib.sb2_obmc_last_kernel() == want
    assert RAN["tma"] > 0 and RAN["scatter"] > 0 and (obmc_kernel == "auto" or RAN[obmc_kernel] > 0)
