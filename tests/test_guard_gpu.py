"""GPU: no kernel writes outside the slabs it is given.  Every slab of these tests sits between two
4 KiB guard zones filled with a pattern (and its padding bytes -- row tails beyond the extended
width, the gap to the next picture -- are patterned too); after the kernels ran the guards and the
padding must be untouched."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
GUARD = 4096
PAT = 0xA5


def guarded(dev, torch, layout, count):
    """A PictureSlab whose memory is carved out of a larger patterned buffer."""
    s = dev.PictureSlab.__new__(dev.PictureSlab)
    s.layout, s.count = layout, count
    s.whole = torch.full((GUARD + layout.pitch * count + 256 + GUARD,), PAT, dtype=torch.uint8, device="cuda")
    s.buf = s.whole[GUARD:GUARD + layout.pitch * count + 256]
    s.slab = s._make_slab()
    return s


def owned_mask(layout, count):
    """True for the bytes a kernel may write: the (extended) planes of every picture."""
    m = np.zeros(layout.pitch * count + 256, bool)
    ext, bpp = layout.extension, layout.bpp
    for p in range(count):
        for c, (w, h) in enumerate(layout.comp_sizes):
            row0 = p * layout.pitch + layout.offset[c] - layout.stride[c] * ext - bpp * ext
            width = (w + 2 * ext) * bpp
            for ph in range(4 if layout.upsampled else 1):
                for y in range(h + 2 * ext):
                    a = row0 + y * layout.stride[c] + (layout.stride[c] >> 2) * ph
                    m[a:a + width] = True
    return m


def check(s, what):
    whole = s.whole.cpu().numpy()
    assert (whole[:GUARD] == PAT).all(), what + ": wrote before the slab"
    assert (whole[-GUARD:] == PAT).all(), what + ": wrote after the slab"
    inner = whole[GUARD:-GUARD]
    mask = owned_mask(s.layout, s.count)
    assert (inner[~mask] == PAT).all(), what + ": wrote into padding between planes / pictures"


def test_kernels_stay_inside_their_slabs(cuda):
    import torch
    from schroedinger_b200 import device as dev
    rng = np.random.default_rng(1)
    w, h, count = 176, 144, 2

    def fill(s, lo, hi):
        for p in range(s.count):
            for c, (pw, ph) in enumerate(s.layout.comp_sizes):
                s.upload(p, c, rng.integers(lo, hi, size=(ph, pw)))

    # wavelets (out of place and in place), s16 and s32
    for depth_name in ("s16", "s32"):
        a = guarded(dev, torch, dev.FrameLayout.yuv420(depth_name, w, h), count)
        b = guarded(dev, torch, dev.FrameLayout.yuv420(depth_name, w, h), count)
        fill(a, -200, 200)
        dev.iwt_forward(a, b, 6, 3)
        dev.iwt_inverse(b, a, 6, 3)
        dev.iwt_forward(a, a, 1, 3)
        dev.iwt_inverse(a, a, 1, 3)
        torch.cuda.synchronize()
        check(a, "wavelet " + depth_name); check(b, "wavelet " + depth_name)
    # edge extension, upsample (fused and separate), downsample with border
    up = guarded(dev, torch, dev.FrameLayout.yuv420("u8", w, h, 32, upsampled=True), count)
    fill(up, 0, 256)
    dev.mc_edgeextend(up)
    dev.upsample(up)
    dev.edgeextend_upsample(up)
    lo = guarded(dev, torch, dev.FrameLayout("u8", [(w // 2, h // 2), (w // 4, h // 4), (w // 4, h // 4)], 8), count)
    src = guarded(dev, torch, dev.FrameLayout.yuv420("u8", w, h, 32), count)
    fill(src, 0, 256)
    dev.mc_edgeextend(src)
    dev.downsample(src, lo)
    dev.downsample_edgeextend(src, lo)
    torch.cuda.synchronize()
    check(up, "upsample"); check(lo, "downsample"); check(src, "edge extension")
    # convert / add, dequantise
    s16 = guarded(dev, torch, dev.FrameLayout.yuv420("s16", w, h + 8), count)
    u8 = guarded(dev, torch, dev.FrameLayout.yuv420("u8", w, h), count)
    fill(s16, -300, 300)
    dev.frame_convert(s16, u8)
    dev.frame_add(s16, u8)
    dev.frame_add(s16, u8, subtract=True)
    q = guarded(dev, torch, dev.FrameLayout.yuv420("s16", 192, 160), count)
    fill(q, -5, 6)
    pairs = torch.tensor([[64, 34]] * (3 * 10 * count), dtype=torch.int32, device="cuda").reshape(-1)
    dev.dequantise(q, 3, [1, 1, 1, 1], [1, 1, 1, 1], pairs)
    torch.cuda.synchronize()
    check(s16, "frame add"); check(u8, "frame convert"); check(q, "dequantise")


def test_block_matching_stays_inside(cuda):
    import torch
    from schroedinger_b200 import device as dev
    rng = np.random.default_rng(2)
    w, h, count = 176, 144, 2
    src, ref = helpers.panning_pair(w, h, np.random.default_rng(3))
    nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
    prm = dev.HbmParams(8, 8, nbx, nby, 0, 0, 1, 1)
    levels = 3
    lay0 = dev.FrameLayout.yuv420("u8", w, h, 32)
    s0, r0 = guarded(dev, torch, lay0, count), guarded(dev, torch, lay0, count)
    for p in range(count):
        for c in range(3):
            s0.upload(p, c, src[c]); r0.upload(p, c, ref[c])
    ps = dev.Pyramid(w, h, count, levels, 8, level0=s0)
    pr = dev.Pyramid(w, h, count, levels, 8, level0=r0)
    ps.build(); pr.build()
    n = nbx * nby
    fields = [torch.full((GUARD + count * n * 20 + GUARD,), PAT, dtype=torch.uint8, device="cuda") for _ in range(levels + 1)]
    dev.hbm_scan(prm, ps, pr, 3, [f[GUARD:GUARD + count * n * 20] for f in fields])
    torch.cuda.synchronize()
    check(s0, "pyramid level 0 (source)"); check(r0, "pyramid level 0 (reference)")
    for f in fields:
        a = f.cpu().numpy()
        assert (a[:GUARD] == PAT).all() and (a[-GUARD:] == PAT).all(), "motion field overrun"


@pytest.mark.parametrize("add", [1, 0])
def test_obmc_stays_inside(cuda, add):
    import torch
    from schroedinger_b200 import device as dev
    oracle = helpers.load_oracle()
    case = helpers.ObmcCase(oracle, rng=np.random.default_rng(4), width=176, height=144, span=200, outliers=0.05)
    sizes, count = case.comp_sizes, 2
    ref_lay = dev.FrameLayout("u8", sizes, 32, True)
    refs = []
    for planes in (case.ref0, case.ref1):
        slab = guarded(dev, torch, ref_lay, count)
        for p in range(count):
            for c, pl in enumerate(planes):
                start = p * ref_lay.pitch + ref_lay.offset[c] - pl.origin
                slab.buf[start:start + pl.buf.size].copy_(torch.from_numpy(pl.buf.reshape(-1)))
        refs.append(slab)
    before = [r.whole.clone() for r in refs]
    res = guarded(dev, torch, dev.FrameLayout("s16", sizes), count)
    acc = guarded(dev, torch, dev.FrameLayout("s16", sizes), count)
    out = guarded(dev, torch, dev.FrameLayout("u8", sizes), count)
    for p in range(count):
        for c in range(3):
            res.upload(p, c, case.residual[c])
    mvs = torch.from_numpy(np.tile(case.mvs.view(np.uint8), count)).cuda()
    prm = dev.ObmcParams(case.xbsep, case.ybsep, case.xblen, case.yblen, case.nbx, case.nby,
                         case.prec, case.weights[0], case.weights[1], case.weights[2], case.hs, case.vs)
    dev.obmc_render(prm, mvs, refs[0], refs[1], res, add, out=out, acc=acc)
    torch.cuda.synchronize()
    check(res, "OBMC residual"); check(acc, "OBMC accumulator"); check(out, "OBMC output")
    for r, b in zip(refs, before):
        assert torch.equal(r.whole, b), "OBMC wrote into a reference frame"


def test_second_round_kernels_stay_inside_their_buffers(cuda):
    """The rows added in round 2: low-delay slice decoder, fused inverse + convert, fused two-level inverse, rough
    search, sub-pel refinement -- slabs between guard zones, motion fields inside patterned tensors."""
    import torch
    from schroedinger_b200 import device as dev, lib
    rng = np.random.default_rng(2)
    # low-delay slices -> coefficient slab; fused inverse + convert -> u8 slab with a border
    w, h, depth, nh, nv, nbytes = 256, 128, 3, 8, 4, 70
    dq = np.load(__import__("os").path.join(helpers.GOLDEN_DIR, "dequant.npz"))
    data = rng.integers(0, 256, size=2 * nh * nv * nbytes, dtype=np.uint8)
    for k in range(2 * nh * nv):                           # declared luma lengths inside the slice
        data[k * nbytes + 1] &= 0x03
    slices = torch.from_numpy(data).cuda()
    for name in ("s16", "s32"):
        coeffs = guarded(dev, torch, dev.FrameLayout.yuv420(name, w, h), 2)
        dev.lowdelay_decode(slices, nh * nv * nbytes, coeffs, depth, nh, nv, nbytes, 1, [0] * (1 + 3 * depth),
                            dq["table_quant"], dq["table_offset_1_2"])
        pic = guarded(dev, torch, dev.FrameLayout.yuv420("u8", w - 6, h - 10, 16), 2)
        dev.iwt_inverse_convert(coeffs, pic, 0, depth, 1)
        out = guarded(dev, torch, dev.FrameLayout.yuv420(name, w, h), 2)
        lib.sb2_iwt_enable_fused(1)
        try:
            big = guarded(dev, torch, dev.FrameLayout.yuv420(name, 640, 384), 1)
            big2 = guarded(dev, torch, dev.FrameLayout.yuv420(name, 640, 384), 1)
            dev.iwt_inverse(big, big2, 6, 2)
        finally:
            lib.sb2_iwt_enable_fused(0)
        dev.iwt_inverse(coeffs, out, 0, depth)
        torch.cuda.synchronize()
        for s_, what in ((coeffs, "lowdelay"), (out, "inverse"), (big2, "fused two-level inverse")):
            check(s_, what + " " + name)
        # the u8 picture: only picture pixels may change (its border is not written either)
        whole = pic.whole.cpu().numpy()
        assert (whole[:GUARD] == PAT).all() and (whole[-GUARD:] == PAT).all(), "inverse + convert wrote outside the slab"
        L = pic.layout
        m = np.zeros(L.pitch * 2 + 256, bool)
        for p in range(2):
            for c, (pw, ph) in enumerate(L.comp_sizes):
                for y in range(ph):
                    a = p * L.pitch + L.offset[c] + y * L.stride[c]
                    m[a:a + pw] = True
        assert (whole[GUARD:-GUARD][~m] == PAT).all(), "inverse + convert wrote outside the picture"
    # rough search + sub-pel refinement: the fields
    w, h, levels, count = 200, 104, 2, 2                   # ragged: partial blocks, blocks outside the coarse levels
    ps, pr = dev.Pyramid(w, h, count, levels, 8), dev.Pyramid(w, h, count, levels, 8)
    for p in range(count):
        s, r = helpers.panning_pair(w, h, rng, (3, -2))
        for c in range(3):
            ps.slabs[0].upload(p, c, s[c])
            pr.slabs[0].upload(p, c, r[c])
    ps.build()
    pr.build()
    nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
    n = nbx * nby * 20 * count
    prm = dev.HbmParams(8, 8, nbx, nby, 1, 0, 1, 1)

    def field():
        t = torch.full((GUARD + n + GUARD,), PAT, dtype=torch.uint8, device="cuda")
        return t, t[GUARD:GUARD + n]
    wholes, fields = [None], [None]
    for _ in range(levels):
        t, f = field()
        wholes.append(t)
        fields.append(f)
    dev.rough_scan(prm, ps, pr, 12, 4, fields=fields)
    up = dev.PictureSlab(dev.FrameLayout.yuv420("u8", w, h, 32, True), count)
    for p in range(count):
        for c in range(3):
            up.upload(p, c, pr.slabs[0].download(p, c))
    dev.edgeextend_upsample(up)
    t0, f0 = field()
    f0.zero_()
    dev.subpel_refine(ps.slabs[0], up, f0, 8, 8, nbx, nby, 3, 1, 0.2)
    # the split-2 pass of the mode decision on that field: decided blocks and the superblock sums
    tm, fm = field()
    nsb = (nbx // 4) * (nby // 4) * count
    te = torch.full((GUARD // 4 + nsb + GUARD // 4,), 0x25A5A5A5, dtype=torch.int32, device="cuda")
    tn = te.clone()
    dev.split2_decide(ps.slabs[0], [up], [f0], 8, 8, nbx, nby, 3, 0.2,
                      out=(fm, te[GUARD // 4:GUARD // 4 + nsb], tn[GUARD // 4:GUARD // 4 + nsb]))
    torch.cuda.synchronize()
    for t in wholes[1:] + [t0, tm]:
        a = t.cpu().numpy()
        assert (a[:GUARD] == PAT).all() and (a[-GUARD:] == PAT).all(), "a motion-field kernel wrote outside its field"
    for t in (te, tn):
        a = t.cpu().numpy()
        assert (a[:GUARD // 4] == 0x25A5A5A5).all() and (a[-GUARD // 4:] == 0x25A5A5A5).all(), "superblock sums overran"
        assert (a[GUARD // 4:-GUARD // 4] != 0x25A5A5A5).all(), "a superblock sum was never written"
