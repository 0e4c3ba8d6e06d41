"""Shared test helpers: ctypes access to the oracle (oracle/liboracle.so), to the
compiled reference (oracle/_ref/libschro_ref.so, optional) and synthetic patterns.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may touch oracle/.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libschro_ref.so")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

FILTER_NAMES = ["desl_dubuc_9_7", "le_gall_5_3", "desl_dubuc_13_7", "haar0", "haar1",
                "fidelity", "daub_9_7"]


def load_oracle():
    if not os.path.exists(ORACLE_PATH):
        import subprocess
        subprocess.check_call(["make", "-C", ROOT, "oracle/liboracle.so"])
    return ctypes.CDLL(ORACLE_PATH, mode=ctypes.RTLD_LOCAL)


def load_ref():
    """The unmodified reference compiled by oracle/build_ref.sh, or None."""
    if not os.path.exists(REF_PATH):
        return None
    return ctypes.CDLL(REF_PATH, mode=ctypes.RTLD_LOCAL)


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def cpu_wavelet(lib, prefix, direction, a, filt, depth=None):
    """Run {oracle,ref}_{wavelet,iwt}_{fwd,inv} in place on a 2-D int16/int32 array."""
    assert a.dtype in (np.int16, np.int32) and a.flags.c_contiguous
    is32 = 1 if a.dtype == np.int32 else 0
    h, w = a.shape
    if depth is None:
        fn = getattr(lib, f"{prefix}_wavelet_{direction}")
        fn.restype = None
        fn(_vp(a), ctypes.c_int(a.strides[0]), w, h, is32, filt)
    else:
        fn = getattr(lib, f"{prefix}_iwt_{direction}")
        fn.restype = None
        fn(_vp(a), ctypes.c_int(a.strides[0]), w, h, is32, filt, depth)
    return a


def lcg(seed, n):
    """The survey's synthetic source: x = 1103515245*x + 12345, take x>>16."""
    out = np.empty(n, dtype=np.uint32)
    x = np.uint64(seed)
    a, c, m = np.uint64(1103515245), np.uint64(12345), np.uint64(0xFFFFFFFF)
    # vectorised in blocks via jump-ahead would be overkill; numpy loop in chunks
    xs = int(seed)
    for i in range(n):
        xs = (1103515245 * xs + 12345) & 0xFFFFFFFF
        out[i] = xs >> 16
    return out


def patterns(h, w, dtype, rng, amp=255):
    """Deterministic test patterns in the spirit of the reference's testsuite/common.c:357-396
    (random, constants, lines, bands, edges, ramps)."""
    pats = []
    pats.append(("random", rng.integers(-amp, amp + 1, size=(h, w))))
    for v in (0, 1, -1, amp, -amp, amp // 2, 100):
        pats.append((f"const{v}", np.full((h, w), v)))
    yy, xx = np.mgrid[0:h, 0:w]
    for k in (1, 2, 3, 4, 8):
        pats.append((f"vlines{k}", np.where(xx % (2 * k) < k, amp, 0)))
        pats.append((f"hlines{k}", np.where(yy % (2 * k) < k, amp, 0)))
    pats.append(("checker", np.where((xx + yy) & 1, amp, -amp)))
    for frac in (0.25, 0.5, 0.75):
        pats.append((f"vedge{frac}", np.where(xx < w * frac, amp, 0)))
        pats.append((f"hedge{frac}", np.where(yy < h * frac, amp, 0)))
    pats.append(("hramp", (xx * amp) // max(1, w - 1)))
    pats.append(("vramp", (yy * amp) // max(1, h - 1)))
    pats.append(("dramp", ((xx + yy) * amp) // max(1, w + h - 2)))
    pats.append(("impulse", np.where((xx == w // 2) & (yy == h // 2), amp, 0)))
    return [(name, np.ascontiguousarray(p.astype(dtype))) for name, p in pats]


# ---------------------------------------------------------------------------------------
# frames, upsampled references, motion fields
# ---------------------------------------------------------------------------------------
MV_DTYPE = np.dtype([("flags", "<u4"), ("metric", "<u4"), ("chroma_metric", "<u4"),
                     ("v", "<i2", (4,))])


def round_up(x, a):
    return (x + a - 1) // a * a


class HostPlane:
    """One u8 component with `ext` border pixels, optionally in the reference's 4-phase
    "upsampled" row layout (schroedinger/schroframe.c:133-137, 1917-1925)."""

    def __init__(self, width, height, ext=0, upsampled=False, fill=0):
        self.w, self.h, self.ext, self.upsampled = width, height, ext, upsampled
        self.stride = round_up(width + 2 * ext, 16) * (4 if upsampled else 1)
        self.buf = np.full((height + 2 * ext, self.stride), fill, dtype=np.uint8)
        self.origin = ext * self.stride + ext            # byte offset of pixel (0,0) of phase 0

    @property
    def ptr(self):
        return ctypes.c_void_p(self.buf.ctypes.data + self.origin)

    def phase(self, p=0, with_border=True):
        q = (self.stride >> 2) * p if self.upsampled else 0
        e = self.ext
        if with_border:
            return self.buf[:, q:q + self.w + 2 * e]
        return self.buf[e:e + self.h, q + e:q + e + self.w]

    def set_image(self, img):
        self.phase(0, with_border=False)[...] = img


def cpu_edgeextend(lib, prefix, plane):
    fn = getattr(lib, f"{prefix}_mc_edgeextend")
    fn.restype = None
    fn(plane.ptr, plane.stride, plane.w, plane.h, plane.ext)


def cpu_upsample(lib, prefix, plane):
    fn = getattr(lib, f"{prefix}_upsample")
    fn.restype = None
    fn(plane.ptr, plane.stride, plane.w, plane.h, plane.ext)


def cpu_downsample(lib, prefix, src):
    """src: 2-D u8 array -> half-size array (schro_frame_downsample on one component)."""
    h, w = src.shape
    dst = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
    fn = getattr(lib, f"{prefix}_downsample")
    fn.restype = None
    fn(dst.ctypes.data_as(ctypes.c_void_p), dst.strides[0], dst.shape[1], dst.shape[0],
       src.ctypes.data_as(ctypes.c_void_p), src.strides[0], w, h)
    return dst


def smooth_image(h, w, rng, noise=5):
    """Smooth gradient + hash noise."""
    yy, xx = np.mgrid[0:h, 0:w]
    base = 128 + 60 * np.sin(xx / 17.0) + 50 * np.cos(yy / 23.0) + 0.05 * (xx + yy)
    img = base + rng.integers(-(1 << noise) // 2, (1 << noise) // 2 + 1, size=(h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


def make_upsampled_ref(oracle, width, height, rng, chroma_shift=(1, 1)):
    """Three upsampled, edge-extended (ext 32) component planes of one reference picture."""
    planes = []
    for k in range(3):
        w = width if k == 0 else (width + (1 << chroma_shift[0]) - 1) >> chroma_shift[0]
        h = height if k == 0 else (height + (1 << chroma_shift[1]) - 1) >> chroma_shift[1]
        pl = HostPlane(w, h, ext=32, upsampled=True)
        pl.set_image(smooth_image(h, w, rng))
        cpu_edgeextend(oracle, "oracle", pl)
        cpu_upsample(oracle, "oracle", pl)
        planes.append(pl)
    return planes


def make_mv_field(nbx, nby, rng, prec=2, span=64, outliers=0.01, modes=(0.12, 0.38, 0.25, 0.25),
                  two_refs=True):
    """pred_mode in {0:12%,1:38%,2:25%,3:25%}, vectors uniform in +-span
    (+1% outliers +-4000), DC uniform +-127."""
    n = nbx * nby
    mv = np.zeros(n, dtype=MV_DTYPE)
    p = np.array(modes, dtype=float)
    if not two_refs:
        p = np.array([p[0], p[1] + p[2] + p[3], 0, 0])
    mode = rng.choice(4, size=n, p=p / p.sum())
    v = rng.integers(-span, span + 1, size=(n, 4))
    out = rng.random(n) < outliers
    v[out] = rng.integers(-4000, 4001, size=(int(out.sum()), 4))
    dc = rng.integers(-127, 128, size=(n, 4))
    dc[:, 3] = 0
    mv["v"] = np.where((mode == 0)[:, None], dc, v)
    mv["flags"] = mode.astype(np.uint32)
    return mv


class ObmcCase:
    """Everything one schro_motion_render call needs, as host arrays."""

    def __init__(self, oracle, width, height, rng, xbsep=8, ybsep=8, xblen=12, yblen=12, prec=2,
                 weights=(1, 1, 1), num_refs=2, chroma_format=2, res_is_s32=False, span=64,
                 outliers=0.01):
        self.width, self.height = width, height
        self.chroma_format = chroma_format
        self.hs = 0 if chroma_format == 0 else 1
        self.vs = 1 if chroma_format == 2 else 0
        self.xbsep, self.ybsep, self.xblen, self.yblen = xbsep, ybsep, xblen, yblen
        self.prec, self.weights, self.num_refs = prec, weights, num_refs
        self.nbx = 4 * ((width + 4 * xbsep - 1) // (4 * xbsep))
        self.nby = 4 * ((height + 4 * ybsep - 1) // (4 * ybsep))
        self.res_is_s32 = res_is_s32
        self.ref0 = make_upsampled_ref(oracle, width, height, rng, (self.hs, self.vs))
        self.ref1 = make_upsampled_ref(oracle, width, height, rng, (self.hs, self.vs)) if num_refs > 1 else None
        self.mvs = make_mv_field(self.nbx, self.nby, rng, prec, span=span, outliers=outliers,
                                 two_refs=num_refs > 1)
        self.comp_sizes = [(p.w, p.h) for p in self.ref0]
        rdt = np.int32 if res_is_s32 else np.int16
        self.residual = [rng.integers(-32, 33, size=(h, w)).astype(rdt) for (w, h) in self.comp_sizes]

    def comp_params(self, k):
        hs, vs = (self.hs, self.vs) if k else (0, 0)
        return dict(xbsep=self.xbsep >> hs, ybsep=self.ybsep >> vs, xblen=self.xblen >> hs,
                    yblen=self.yblen >> vs, x_num_blocks=self.nbx, y_num_blocks=self.nby,
                    mv_precision=self.prec, weight1=self.weights[0], weight2=self.weights[1],
                    weight_bits=self.weights[2], h_shift=hs, v_shift=vs, comp=k)


class OracleObmcParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "xbsep", "ybsep", "xblen", "yblen", "x_num_blocks", "y_num_blocks", "mv_precision",
        "weight1", "weight2", "weight_bits", "h_shift", "v_shift", "comp")]


class RefMotionParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "width", "height", "chroma_format", "xbsep", "ybsep", "xblen", "yblen", "x_num_blocks",
        "y_num_blocks", "mv_precision", "weight1", "weight2", "weight_bits", "num_refs")]


def oracle_obmc(oracle, case, add):
    """-> per component (acc, residual_after, out) from the oracle."""
    res = []
    fn = oracle.oracle_obmc_render
    fn.restype = None
    for k, (w, h) in enumerate(case.comp_sizes):
        p = OracleObmcParams(**case.comp_params(k))
        acc = np.zeros((h, w), np.int16)
        resid = case.residual[k].copy()
        out = np.zeros((h, w), np.uint8)
        r1 = case.ref1[k].ptr if case.ref1 else None
        fn(ctypes.byref(p), case.mvs.ctypes.data_as(ctypes.c_void_p), case.ref0[k].ptr, r1,
           case.ref0[k].stride, w, h, acc.ctypes.data_as(ctypes.c_void_p), acc.strides[0],
           resid.ctypes.data_as(ctypes.c_void_p), resid.strides[0], 1 if case.res_is_s32 else 0,
           1 if add else 0, out.ctypes.data_as(ctypes.c_void_p), out.strides[0])
        res.append((acc, resid, out))
    return res


def oracle_obmc_ref(oracle, case, add, global_motion):
    """The reference's per-pixel renderer (schro_motion_render_ref) through the oracle; global_motion: 20 ints."""
    res = []
    fn = oracle.oracle_obmc_render_ref
    fn.restype = None
    gm = (ctypes.c_int * 20)(*[int(v) for v in global_motion])
    for k, (w, h) in enumerate(case.comp_sizes):
        p = OracleObmcParams(**case.comp_params(k))
        acc = np.zeros((h, w), np.int16)
        resid = case.residual[k].astype(np.int16).copy()
        out = np.zeros((h, w), np.uint8)
        r1 = case.ref1[k].ptr if case.ref1 else None
        fn(ctypes.byref(p), gm, case.mvs.ctypes.data_as(ctypes.c_void_p), case.ref0[k].ptr, r1, case.ref0[k].stride, w, h,
           acc.ctypes.data_as(ctypes.c_void_p), acc.strides[0] // 2, resid.ctypes.data_as(ctypes.c_void_p),
           resid.strides[0] // 2, 1 if add else 0, out.ctypes.data_as(ctypes.c_void_p), out.strides[0])
        res.append((acc, resid, out))
    return res


def ref_obmc_global(ref, case, add, global_motion):
    """schro_motion_render with have_global_motion set through the compiled reference."""
    gm = (ctypes.c_int * 20)(*[int(v) for v in global_motion])
    ref.ref_set_global_motion(gm)
    try:
        return ref_obmc(ref, case, add)
    finally:
        ref.ref_set_global_motion(None)


def global_motion_case(oracle, width, height, rng, **kw):
    """An ObmcCase in which a third of the blocks use the picture's global-motion model, plus that model:
    a slight zoom / rotation / pan per reference (a_exp = 8) with a small perspective term."""
    case = ObmcCase(oracle, width, height, rng, **kw)
    mv = case.mvs
    use = rng.random(len(mv)) < 0.35
    mv["flags"] = np.where(use & ((mv["flags"] & 3) != 0), mv["flags"] | 4, mv["flags"])
    gm = []
    for r in range(2):
        # b0 b1 a_exp a00 a01 a10 a11 c_exp c0 c1: the vector is the model's displacement, ((A x + 2^a b) * scale) >> (a + c)
        gm += [int(rng.integers(-6, 7)), int(rng.integers(-6, 7)), 8, int(rng.integers(-3, 4)), int(rng.integers(-2, 3)),
               int(rng.integers(-2, 3)), int(rng.integers(-3, 4)), 12, int(rng.integers(-1, 2)), int(rng.integers(-1, 2))]
    return case, gm


def ref_obmc(ref, case, add, use_ref_renderer=False):
    """Same through the unmodified reference (schro_motion_render_u8 or _ref)."""
    P = ctypes.c_void_p * 3
    I = ctypes.c_int * 3
    accs = [np.zeros((h, w), np.int16) for (w, h) in case.comp_sizes]
    resid = [r.copy() for r in case.residual]
    outs = [np.zeros((h, w), np.uint8) for (w, h) in case.comp_sizes]
    mp = RefMotionParams(case.width, case.height, case.chroma_format, case.xbsep, case.ybsep,
                         case.xblen, case.yblen, case.nbx, case.nby, case.prec, case.weights[0],
                         case.weights[1], case.weights[2], case.num_refs)
    r0 = P(*[p.ptr for p in case.ref0])
    r0s = I(*[p.stride for p in case.ref0])
    r1 = P(*[p.ptr for p in case.ref1]) if case.ref1 else None
    r1s = I(*[p.stride for p in case.ref1]) if case.ref1 else None
    fn = ref.ref_motion_render
    fn.restype = None
    fn(ctypes.byref(mp), case.mvs.ctypes.data_as(ctypes.c_void_p), r0, r0s, r1, r1s,
       P(*[a.ctypes.data for a in accs]), I(*[a.strides[0] for a in accs]),
       P(*[a.ctypes.data for a in resid]), I(*[a.strides[0] for a in resid]),
       1 if case.res_is_s32 else 0, 1 if add else 0,
       P(*[a.ctypes.data for a in outs]), I(*[a.strides[0] for a in outs]),
       1 if use_ref_renderer else 0)
    return list(zip(accs, resid, outs))


# ---------------------------------------------------------------------------------------
# pyramids and hierarchical block matching
# ---------------------------------------------------------------------------------------
class OraclePyrLevel(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p * 3), ("stride", ctypes.c_int * 3), ("width", ctypes.c_int),
                ("height", ctypes.c_int), ("h_shift", ctypes.c_int), ("v_shift", ctypes.c_int),
                ("ext", ctypes.c_int)]


class RefHbmParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("width", "height", "chroma_format", "xbsep", "ybsep",
                                            "levels", "use_chroma", "ref_index", "level0_range")]


def panning_pair(width, height, rng, pan=(5, 3), noise=3):
    """A textured base picture and a copy panned by `pan` pixels + noise;
    returns (src_planes, ref_planes) as lists of three 2-D u8 arrays (4:2:0)."""
    H, W = height + 64, width + 64
    yy, xx = np.mgrid[0:H, 0:W]
    base = (128 + 50 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + 30 * np.sin((xx + 2 * yy) / 31.0)
            + rng.integers(-20, 21, size=(H, W)))

    def crop(ox, oy):
        y = np.clip(base[32 + oy:32 + oy + height, 32 + ox:32 + ox + width]
                    + rng.integers(-noise, noise + 1, size=(height, width)), 0, 255).astype(np.uint8)
        c = y[::2, ::2]
        return [y, np.ascontiguousarray(255 - c), np.ascontiguousarray((c // 2 + 64).astype(np.uint8))]

    return crop(0, 0), crop(pan[0], pan[1])


def build_pyramid(lib, prefix, planes, levels, ext_level0=32, ext=8):
    """List (level 0..levels) of lists of three HostPlanes: level 0 = the picture (ext 32),
    level i+1 = downsample of level i with extension max(xbsep, ybsep), edge-extended
    (schroedinger/schroanalysis.c:9-28)."""
    pyr = []
    cur = []
    for a in planes:
        pl = HostPlane(a.shape[1], a.shape[0], ext=ext_level0)
        pl.set_image(a)
        cpu_edgeextend(lib, prefix, pl)
        cur.append(pl)
    pyr.append(cur)
    for _ in range(levels):
        nxt = []
        for pl in cur:
            d = cpu_downsample(lib, prefix, np.ascontiguousarray(pl.phase(0, with_border=False)))
            q = HostPlane(d.shape[1], d.shape[0], ext=ext)
            q.set_image(d)
            cpu_edgeextend(lib, prefix, q)
            nxt.append(q)
        pyr.append(nxt)
        cur = nxt
    return pyr


def pyr_level_struct(level_planes, hs=1, vs=1):
    s = OraclePyrLevel()
    for k, pl in enumerate(level_planes):
        s.data[k] = pl.buf.ctypes.data + pl.origin
        s.stride[k] = pl.stride
    s.width, s.height = level_planes[0].w, level_planes[0].h
    s.h_shift, s.v_shift = hs, vs
    s.ext = level_planes[0].ext
    return s


def hbm_block_counts(width, height, xbsep, ybsep):
    return (4 * ((width + 4 * xbsep - 1) // (4 * xbsep)), 4 * ((height + 4 * ybsep - 1) // (4 * ybsep)))


def oracle_hbm(oracle, src_planes, ref_planes, width, height, xbsep=8, ybsep=8, levels=4,
               use_chroma=0, ref_index=0, level0_range=3):
    """schro_hbm_scan + optional level-0 refinement through the oracle -> fields[level]."""
    ext = max(xbsep, ybsep)
    ps = build_pyramid(oracle, "oracle", src_planes, levels, ext=ext)
    pr = build_pyramid(oracle, "oracle", ref_planes, levels, ext=ext)
    nbx, nby = hbm_block_counts(width, height, xbsep, ybsep)
    fields = np.zeros((levels + 1, nbx * nby), dtype=MV_DTYPE)
    fn = oracle.oracle_hbm_scan_hint
    fn.restype = None
    rng_ = 20
    order = [(levels, rng_)]
    r = rng_ >> 1
    for l in range(levels - 1, 0, -1):
        order.append((l, max(3, r)))
        r >>= 1
    if level0_range > 0:
        order.append((0, level0_range))
    for (l, hr) in order:
        s, rr = pyr_level_struct(ps[l]), pyr_level_struct(pr[l])
        parent = fields[l + 1].ctypes.data_as(ctypes.c_void_p) if l < levels else None
        fn(ctypes.byref(s), ctypes.byref(rr), xbsep, ybsep, nbx, nby, ref_index, l, hr, use_chroma,
           parent, fields[l].ctypes.data_as(ctypes.c_void_p))
    return fields, ps, pr


def ref_hbm(ref, src_planes, ref_planes, width, height, xbsep=8, ybsep=8, levels=4, use_chroma=0,
            ref_index=0, level0_range=3, scratch=None):
    """schro_hbm_scan + level-0 refinement through the compiled reference.  scratch: a dict that keeps the
    output buffers between calls (bench.py's CPU arm: no allocation inside its timed loop)."""
    nbx, nby = hbm_block_counts(width, height, xbsep, ybsep)
    P = ctypes.c_void_p * 3
    I = ctypes.c_int * 3
    hp = RefHbmParams(width, height, 2, xbsep, ybsep, levels, use_chroma, ref_index, level0_range)
    nx, ny = ctypes.c_int(), ctypes.c_int()
    if scratch is not None and "fields" in scratch:
        fields, pyr, ptrs = scratch["fields"], scratch["pyr"], scratch["ptrs"]
    else:
        fields = np.zeros((levels + 1, nbx * nby), dtype=MV_DTYPE)
        pyr = []
        ptrs = (ctypes.c_void_p * (levels * 3))()
        w, h = width, height
        # luma and chroma sizes follow their own halving chains
        cw, ch = (width + 1) // 2, (height + 1) // 2
        for l in range(levels):
            w, h, cw, ch = (w + 1) // 2, (h + 1) // 2, (cw + 1) // 2, (ch + 1) // 2
            pyr.append([np.zeros((h, w), np.uint8), np.zeros((ch, cw), np.uint8), np.zeros((ch, cw), np.uint8)])
            for k in range(3):
                ptrs[l * 3 + k] = pyr[l][k].ctypes.data
        if scratch is not None:
            scratch.update(fields=fields, pyr=pyr, ptrs=ptrs)
    fn = ref.ref_hbm_run
    fn.restype = None
    fn(ctypes.byref(hp), P(*[a.ctypes.data for a in src_planes]), I(*[a.strides[0] for a in src_planes]),
       P(*[a.ctypes.data for a in ref_planes]), I(*[a.strides[0] for a in ref_planes]),
       fields.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nx), ctypes.byref(ny), ptrs)
    assert (nx.value, ny.value) == (nbx, nby)
    return fields, pyr


def oracle_rough(oracle, src_planes, ref_planes, width, height, xbsep=8, ybsep=8, levels=4, ref_index=0,
                 nohint_distance=12, hint_distance=4):
    """schro_rough_me_heirarchical_scan through the oracle -> fields[level] (level 0 stays zero)."""
    ext = max(xbsep, ybsep)
    ps = build_pyramid(oracle, "oracle", src_planes, levels, ext=ext)
    pr = build_pyramid(oracle, "oracle", ref_planes, levels, ext=ext)
    nbx, nby = hbm_block_counts(width, height, xbsep, ybsep)
    fields = np.zeros((levels + 1, nbx * nby), dtype=MV_DTYPE)
    oracle.oracle_rough_scan_nohint.restype = None
    oracle.oracle_rough_scan_hint.restype = None
    s, rr = pyr_level_struct(ps[levels]), pyr_level_struct(pr[levels])
    oracle.oracle_rough_scan_nohint(ctypes.byref(s), ctypes.byref(rr), xbsep, ybsep, nbx, nby, ref_index, levels,
                                    nohint_distance, fields[levels].ctypes.data_as(ctypes.c_void_p))
    for l in range(levels - 1, 0, -1):
        s, rr = pyr_level_struct(ps[l]), pyr_level_struct(pr[l])
        oracle.oracle_rough_scan_hint(ctypes.byref(s), ctypes.byref(rr), xbsep, ybsep, nbx, nby, ref_index, l,
                                      hint_distance, fields[l + 1].ctypes.data_as(ctypes.c_void_p),
                                      fields[l].ctypes.data_as(ctypes.c_void_p))
    return fields, ps, pr


def ref_rough(ref, src_planes, ref_planes, width, height, xbsep=8, ybsep=8, levels=4, ref_index=0,
              nohint_distance=12, hint_distance=4):
    """The same through the compiled reference (schro_rough_me_heirarchical_scan when the distances
    are the reference's own 12 / 4)."""
    nbx, nby = hbm_block_counts(width, height, xbsep, ybsep)
    P = ctypes.c_void_p * 3
    I = ctypes.c_int * 3
    hp = RefHbmParams(width, height, 2, xbsep, ybsep, levels, 0, ref_index, 0)
    nx, ny = ctypes.c_int(), ctypes.c_int()
    fields = np.zeros((levels + 1, nbx * nby), dtype=MV_DTYPE)
    fn = ref.ref_rough_run
    fn.restype = None
    fn(ctypes.byref(hp), nohint_distance, hint_distance,
       P(*[a.ctypes.data for a in src_planes]), I(*[a.strides[0] for a in src_planes]),
       P(*[a.ctypes.data for a in ref_planes]), I(*[a.strides[0] for a in ref_planes]),
       fields.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nx), ctypes.byref(ny))
    assert (nx.value, ny.value) == (nbx, nby)
    return fields


def rough_inside_mask(width, height, xbsep, ybsep, nbx, nby, level):
    """Blocks of `level`'s grid that overlap that level's frame (the others are where the
    reference's result is undefined, oracle_rough.c)."""
    w, h = width, height
    for _ in range(level):
        w, h = (w + 1) // 2, (h + 1) // 2
    skip = 1 << level
    m = np.zeros((nby, nbx), dtype=bool)
    for j in range(0, nby, skip):
        for i in range(0, nbx, skip):
            m[j, i] = (i >> level) * xbsep < w and (j >> level) * ybsep < h
    return m.reshape(-1)


def subpel_case(oracle, width, height, rng, pans=((5, 3), (-4, 2)), levels=2, num_refs=2):
    """A source picture, `num_refs` panned references and, per reference, the level-0 field of
    hierarchical block matching (what schro_encoder's deep estimation hands to the sub-pel refinement,
    schroedinger/schroencoder.c:2302-2320).  Returns (src_planes, [ref_planes], [field])."""
    src, refs, fields = None, [], []
    base_rng = np.random.default_rng(int(rng.integers(1 << 30)))
    state = base_rng.bit_generator.state
    for r in range(num_refs):
        base_rng.bit_generator.state = state           # the same base picture for every reference
        s, rf = panning_pair(width, height, base_rng, pans[r])
        if src is None:
            src = s
        f, _, _ = oracle_hbm(oracle, src, rf, width, height, levels=levels, ref_index=r)
        refs.append(rf)
        fields.append(f[0].copy())
    return src, refs, fields


def _luma_up(oracle, plane_img):
    pl = HostPlane(plane_img.shape[1], plane_img.shape[0], ext=32, upsampled=True)
    pl.set_image(plane_img)
    cpu_edgeextend(oracle, "oracle", pl)
    cpu_upsample(oracle, "oracle", pl)
    return pl


def oracle_subpel(oracle, src, refs, fields, width, height, xblen=8, yblen=8, mv_precision=2, lam=0.1, orig_ext=32):
    """schro_encoder_motion_predict_subpel_deep through the oracle; returns the refined fields."""
    nbx, nby = hbm_block_counts(width, height, xblen, yblen)
    fn = oracle.oracle_subpel_refine
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                   ctypes.c_void_p]
    y = np.ascontiguousarray(src[0])
    out = []
    for r, rf in enumerate(refs):
        up = _luma_up(oracle, rf[0])
        f = fields[r].copy()
        fn(y.ctypes.data, y.strides[0], width, height, orig_ext, up.buf.ctypes.data + up.origin, up.stride,
           xblen, yblen, nbx, nby, mv_precision, r, lam, f.ctypes.data)
        out.append(f)
    return out


def ref_subpel(ref, src, refs, fields, width, height, xblen=8, yblen=8, mv_precision=2, lam=0.1):
    """The same through the compiled reference (a SchroMe built by its own schro_me_new)."""
    nbx, nby = hbm_block_counts(width, height, xblen, yblen)
    P = ctypes.c_void_p * 3
    I = ctypes.c_int * 3
    q = (ctypes.c_int * 6)(width, height, xblen, yblen, mv_precision, len(refs))
    out = [f.copy() for f in fields]
    nx, ny = ctypes.c_int(), ctypes.c_int()
    fn = ref.ref_subpel_run
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_double] + [ctypes.c_void_p] * 10
    r1 = refs[1] if len(refs) > 1 else refs[0]
    fn(q, lam, P(*[a.ctypes.data for a in src]), I(*[a.strides[0] for a in src]),
       P(*[a.ctypes.data for a in refs[0]]), I(*[a.strides[0] for a in refs[0]]),
       P(*[a.ctypes.data for a in r1]), I(*[a.strides[0] for a in r1]),
       out[0].ctypes.data, out[1].ctypes.data if len(out) > 1 else None, ctypes.byref(nx), ctypes.byref(ny))
    assert (nx.value, ny.value) == (nbx, nby)
    return out


# ---- split-2 pass of the mode decision (schroedinger/schromotionest.c:1601-1802) ------------------------
REF_ME_PATH = os.path.join(ROOT, "oracle", "_ref", "libschro_ref_me.so")


def load_ref_me():
    """schromotionest.c's file-static functions (oracle/ref_me_static.c), or None."""
    if not os.path.exists(REF_ME_PATH) or not os.path.exists(REF_PATH):
        return None
    # RTLD_LOCAL: libschro_ref.so comes in as its dependency (rpath $ORIGIN); making the reference's schro_*
    # symbols global would capture the bindings of everything loaded later (the compat shims, the product library)
    return ctypes.CDLL(REF_ME_PATH, mode=ctypes.RTLD_LOCAL)


class OracleSplit2Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("width", "height", "h_shift", "v_shift", "orig_ext", "xblen", "yblen",
                                            "x_num_blocks", "y_num_blocks", "mv_precision", "num_refs")] + [("lam", ctypes.c_double)]


def split2_case(oracle, width, height, rng, mv_precision=2, num_refs=2, pans=((5, 3), (-4, 2)), lam=0.1):
    """Source, references and the references' sub-pel fields (block matching, then the sub-pel refinement:
    what schro_encoder_mode_decision copies into split2_mf, schroedinger/schroencoder.c:2340-2349)."""
    src, refs, fields = subpel_case(oracle, width, height, rng, pans=pans, num_refs=num_refs)
    if mv_precision > 0:
        fields = oracle_subpel(oracle, src, refs, fields, width, height, 8, 8, mv_precision, lam)
    return src, refs, fields


def oracle_split2(oracle, src, refs, fields, width, height, xblen=8, yblen=8, mv_precision=2, lam=0.1, orig_ext=32):
    """-> (motion, sb_error, sb_entropy) through oracle_split2_decide"""
    nbx, nby = hbm_block_counts(width, height, xblen, yblen)
    p = OracleSplit2Params(width, height, 1, 1, orig_ext, xblen, yblen, nbx, nby, mv_precision, len(refs), lam)
    P = ctypes.c_void_p * 3
    I = ctypes.c_int * 3
    srcs = [np.ascontiguousarray(a) for a in src]
    ups = [[_luma_up(oracle, a) for a in rf] for rf in refs]
    motion = np.zeros(nbx * nby, MV_DTYPE)
    nsb = (nbx // 4) * (nby // 4)
    sb_error, sb_entropy = np.zeros(nsb, np.int32), np.zeros(nsb, np.int32)
    up1 = ups[1] if len(ups) > 1 else ups[0]
    f1 = fields[1] if len(fields) > 1 else fields[0]
    fn = oracle.oracle_split2_decide
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p] * 11
    fn(ctypes.byref(p), P(*[a.ctypes.data for a in srcs]), I(*[a.strides[0] for a in srcs]),
       P(*[u.buf.ctypes.data + u.origin for u in ups[0]]), P(*[u.buf.ctypes.data + u.origin for u in up1]),
       I(*[u.stride for u in ups[0]]), fields[0].ctypes.data, f1.ctypes.data, motion.ctypes.data,
       sb_error.ctypes.data, sb_entropy.ctypes.data)
    return motion, sb_error, sb_entropy


def ref_split2(ref_me, src, refs, fields, width, height, xblen=8, yblen=8, mv_precision=2, lam=0.1):
    """The same through the compiled reference's schro_do_split2 (oracle/ref_me_static.c)."""
    nbx, nby = hbm_block_counts(width, height, xblen, yblen)
    P = ctypes.c_void_p * 3
    I = ctypes.c_int * 3
    q = (ctypes.c_int * 6)(width, height, xblen, yblen, mv_precision, len(refs))
    f = [x.copy() for x in fields]
    motion = np.zeros(nbx * nby, MV_DTYPE)
    nsb = (nbx // 4) * (nby // 4)
    sb_error, sb_entropy = np.zeros(nsb, np.int32), np.zeros(nsb, np.int32)
    nx, ny = ctypes.c_int(), ctypes.c_int()
    r1 = refs[1] if len(refs) > 1 else refs[0]
    fn = ref_me.ref_split2_pass
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_double] + [ctypes.c_void_p] * 13
    fn(q, lam, P(*[a.ctypes.data for a in src]), I(*[a.strides[0] for a in src]),
       P(*[a.ctypes.data for a in refs[0]]), I(*[a.strides[0] for a in refs[0]]),
       P(*[a.ctypes.data for a in r1]), I(*[a.strides[0] for a in r1]),
       f[0].ctypes.data, f[1].ctypes.data if len(f) > 1 else None, motion.ctypes.data,
       sb_error.ctypes.data, sb_entropy.ctypes.data, ctypes.byref(nx), ctypes.byref(ny))
    assert (nx.value, ny.value) == (nbx, nby)
    return motion, sb_error, sb_entropy


# ---- low-delay slices (schroedinger/schrolowdelay.c) -----------------------------------------------
class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value, n):
        self.bits += [(value >> (n - 1 - i)) & 1 for i in range(n)]

    def sint(self, v):
        """interleaved exp-Golomb, schro_pack_encode_sint (schroedinger/schropack.c:149-180)"""
        a = abs(int(v)) + 1
        n = a.bit_length()
        for i in range(n - 1):
            self.bits += [0, (a >> (n - 2 - i)) & 1]
        self.bits.append(1)
        if v:
            self.bits.append(1 if v < 0 else 0)


def ilog2up(x):
    return int(x).bit_length()


def lowdelay_slice_sizes(num, denom, nh, nv):
    n_bytes, rem = num // denom, num % denom
    acc, sizes = 0, []
    for _ in range(nh * nv):
        acc += rem
        extra = 0
        if acc >= denom:
            extra, acc = 1, acc - denom
        sizes.append(n_bytes + extra)
    return sizes


def lowdelay_band_blocks(width, height, depth, nh, nv, sx, sy):
    """(row slice, column slice) of every subband's codeblock of slice (sx, sy) inside a width x height
    coefficient plane in the in-place layout (schro_subband_get_frame_data + schro_frame_data_get_codeblock)."""
    out = []
    for index in range(1 + 3 * depth):
        level = 0 if index == 0 else (index - 1) // 3
        orient = 0 if index == 0 else (index - 1) % 3 + 1
        shift = depth - level
        bw, bh = width >> shift, height >> shift
        step = 1 << shift
        x0, x1 = bw * sx // nh, bw * (sx + 1) // nh
        y0, y1 = bh * sy // nv, bh * (sy + 1) // nv
        ry = [y * step + (step >> 1 if orient & 2 else 0) for y in range(y0, y1)]
        cx = [x + (bw if orient & 1 else 0) for x in range(x0, x1)]
        out.append((ry, cx))
    return out


def lowdelay_encode(quantised, depth, nh, nv, num, denom, rng, base_indices=None, truncate=0.0, fast_lengths=False):
    """Packs quantised coefficient planes (three 2-D int arrays, in-place subband layout) into low-delay
    slices.  Coefficients that do not fit a slice are dropped from the end (the decoder then reads 1 bits =
    zeros / negative signs, schrounpack.c:96-103); `truncate` > 0 cuts the declared luma length of that
    fraction of the slices short on purpose.  Returns (bytes, base_index per slice)."""
    sizes = lowdelay_slice_sizes(num, denom, nh, nv)
    n_bytes = num // denom
    out = bytearray()
    bases = []
    k = 0
    for sy in range(nv):
        for sx in range(nh):
            nbytes = sizes[k]
            total = 8 * nbytes
            base = int(base_indices[k]) if base_indices is not None else int(rng.integers(0, 60))    # the fast path has tables for base < 60 only (schrolowdelay.c:473)
            bases.append(base)
            lb = ilog2up(8 * (n_bytes if fast_lengths else nbytes))
            yw, uw = BitWriter(), BitWriter()
            for (ry, cx) in lowdelay_band_blocks(quantised[0].shape[1], quantised[0].shape[0], depth, nh, nv, sx, sy):
                for y in ry:
                    for x in cx:
                        yw.sint(quantised[0][y, x])
            for (ry, cx) in lowdelay_band_blocks(quantised[1].shape[1], quantised[1].shape[0], depth, nh, nv, sx, sy):
                for y in ry:
                    for x in cx:
                        uw.sint(quantised[1][y, x])
                        uw.sint(quantised[2][y, x])
            room = total - 7 - lb
            ylen = min(len(yw.bits), room, (1 << lb) - 1)
            if truncate and rng.random() < truncate:
                ylen = int(rng.integers(0, ylen + 1))
            w = BitWriter()
            w.put(base, 7)
            w.put(ylen, lb)
            w.bits += yw.bits[:ylen]
            w.bits += uw.bits[:max(0, room - ylen)]
            w.bits += [int(b) for b in rng.integers(0, 2, size=max(0, total - len(w.bits)))]     # junk padding
            bits = np.array(w.bits[:total], dtype=np.uint8)
            out += np.packbits(bits).tobytes()
            k += 1
    return bytes(out), bases


def cpu_lowdelay(lib, prefix, data, width, height, depth, nh, nv, num, denom, quant_matrix, is_s32, path=0, tables=None):
    """Decode through the oracle (prefix "oracle"; path != 0 selects the 16-bit dequantiser of the s16 fast
    path) or the compiled reference (prefix "ref"; path 0 dispatcher, 1 _slow, 2 _fast).  Returns three planes."""
    dt = np.int32 if is_s32 else np.int16
    planes = [np.full((height, width), -1, dt), np.full((height // 2, width // 2), -1, dt), np.full((height // 2, width // 2), -1, dt)]
    P = ctypes.c_void_p * 3
    I = ctypes.c_int * 3
    qm = (ctypes.c_int * len(quant_matrix))(*quant_matrix)
    buf = (ctypes.c_uint8 * (len(data) + 8)).from_buffer_copy(data + b"\0" * 8)
    if prefix == "ref":
        q = (ctypes.c_int * 9)(width, height, depth, nh, nv, num, denom, int(is_s32), path)
        fn = lib.ref_lowdelay_decode
        fn.restype = None
        fn(q, qm, buf, len(data), P(*[a.ctypes.data for a in planes]), I(*[a.strides[0] for a in planes]))
    else:
        tq, to, _ = tables
        fn = lib.oracle_lowdelay_decode
        fn.restype = None
        fn(buf, len(data), num, denom, nh, nv, depth, qm, tq.ctypes.data_as(ctypes.c_void_p), to.ctypes.data_as(ctypes.c_void_p),
           P(*[a.ctypes.data for a in planes]), I(*[a.strides[0] // a.itemsize for a in planes]),
           I(width, width // 2, width // 2), I(height, height // 2, height // 2), int(is_s32), int(path))
    return planes


# ---- combine / convert glue --------------------------------------------
DEPTH_DTYPE = {0: np.uint8, 1: np.int16, 2: np.int32}


def chroma_size(w, h):
    return (w + 1) // 2, (h + 1) // 2


def random_planes(rng, depth, w, h, full_range=True):
    """Three 4:2:0 planes of the given depth; full_range exercises every wrap / saturation point."""
    cw, ch = chroma_size(w, h)
    out = []
    for (pw, ph) in ((w, h), (cw, ch), (cw, ch)):
        if depth == 0:
            a = rng.integers(0, 256, size=(ph, pw))
        elif depth == 1:
            a = rng.integers(-32768, 32768, size=(ph, pw)) if full_range else rng.integers(-600, 600, size=(ph, pw))
        else:
            a = rng.integers(-2 ** 31, 2 ** 31, size=(ph, pw)) if full_range else rng.integers(-70000, 70000, size=(ph, pw))
            a[::3, ::5] = rng.integers(-400, 400, size=a[::3, ::5].shape)      # values near the u8 range too
        out.append(np.ascontiguousarray(a.astype(DEPTH_DTYPE[depth])))
    return out


def oracle_convert(lib, src_planes, sdepth, sw, sh, ddepth, dw, dh):
    lib.oracle_convert_plane.restype = None
    cw, ch = chroma_size(dw, dh)
    scw, sch = chroma_size(sw, sh)
    out = []
    for k, (pw, ph, spw, sph) in enumerate(((dw, dh, sw, sh), (cw, ch, scw, sch), (cw, ch, scw, sch))):
        d = np.zeros((ph, pw), DEPTH_DTYPE[ddepth])
        s = src_planes[k]
        lib.oracle_convert_plane(_vp(d), ctypes.c_int(d.strides[0]), ddepth, pw, ph,
                                 _vp(s), ctypes.c_int(s.strides[0]), sdepth, spw, sph)
        out.append(d)
    return out


def ref_convert(lib, src_planes, sdepth, sw, sh, ddepth, dw, dh):
    lib.ref_frame_convert.restype = None
    cw, ch = chroma_size(dw, dh)
    dst = [np.zeros((dh, dw), DEPTH_DTYPE[ddepth]), np.zeros((ch, cw), DEPTH_DTYPE[ddepth]),
           np.zeros((ch, cw), DEPTH_DTYPE[ddepth])]
    P3, I3 = ctypes.c_void_p * 3, ctypes.c_int * 3
    lib.ref_frame_convert(P3(*[a.ctypes.data for a in dst]), I3(*[a.strides[0] for a in dst]), ddepth, dw, dh,
                          P3(*[a.ctypes.data for a in src_planes]), I3(*[a.strides[0] for a in src_planes]),
                          sdepth, sw, sh)
    return dst


def cpu_add(lib, prefix, dst_planes, dw, dh, src_planes, sdepth, sw, sh, subtract):
    """{oracle_add_plane per plane, ref_frame_add per frame}: dst (s16) +-= src, in place on copies."""
    dst = [a.copy() for a in dst_planes]
    if prefix == "ref":
        lib.ref_frame_add.restype = None
        P3, I3 = ctypes.c_void_p * 3, ctypes.c_int * 3
        lib.ref_frame_add(P3(*[a.ctypes.data for a in dst]), I3(*[a.strides[0] for a in dst]), dw, dh,
                          P3(*[a.ctypes.data for a in src_planes]), I3(*[a.strides[0] for a in src_planes]),
                          sdepth, sw, sh, int(subtract))
        return dst
    lib.oracle_add_plane.restype = None
    for k in range(3):
        d, s = dst[k], src_planes[k]
        lib.oracle_add_plane(_vp(d), ctypes.c_int(d.strides[0]), d.shape[1], d.shape[0],
                             _vp(s), ctypes.c_int(s.strides[0]), sdepth, s.shape[1], s.shape[0], int(subtract))
    return dst


# ---- dequantisation ---------------------------------------------------
def dequant_table_size(depth, hcb, vcb):
    """(factor, offset) pairs for one component: band index 0..3*depth, codeblock rows, columns."""
    n = hcb[0] * vcb[0]
    for level in range(depth):
        n += 3 * hcb[level + 1] * vcb[level + 1]
    return n


def cpu_dequantise(lib, prefix, plane, depth, hcb, vcb, quant):
    """{oracle,ref}_dequantise_plane in place on a copy; quant: int32 array of (factor, offset+2) pairs."""
    a = np.ascontiguousarray(plane.copy())
    fn = getattr(lib, f"{prefix}_dequantise_plane")
    fn.restype = None
    I8 = ctypes.c_int * 8
    h = list(hcb) + [1] * (8 - len(hcb))
    v = list(vcb) + [1] * (8 - len(vcb))
    q = np.ascontiguousarray(quant, dtype=np.int32)
    fn(_vp(a), ctypes.c_int(a.strides[0]), a.shape[1], a.shape[0], 1 if a.dtype == np.int32 else 0, depth,
       I8(*h), I8(*v), _vp(q))
    return a


def ref_quant_tables(lib):
    f, a, b = (np.zeros(61, np.uint32) for _ in range(3))
    lib.ref_quant_tables.restype = None
    lib.ref_quant_tables(_vp(f), _vp(a), _vp(b))
    return f, a, b
