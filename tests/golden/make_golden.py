#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED reference
(oracle/_ref/libschro_ref.so, built by oracle/build_ref.sh from /root/reference).

Run in the build container (where /root/reference exists):
    make ref && python tests/golden/make_golden.py
The fixtures are small .npz files; the GPU box only reads them.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from tests import helpers  # noqa: E402


def wavelet_golden(ref):
    rng = np.random.default_rng(20261018)
    out = {}
    for dtype, tag, amps in ((np.int16, "s16", (255, 32767)), (np.int32, "s32", (1023, 2 ** 31 - 1))):
        for filt in range(7):
            for (h, w) in ((20, 20), (6, 34), (2, 2), (48, 18)):
                for amp in amps:
                    a = rng.integers(-amp, amp + 1, size=(h, w)).astype(dtype)
                    key = f"{tag}_f{filt}_{h}x{w}_a{amp}"
                    out[key + "_in"] = a
                    out[key + "_fwd"] = helpers.cpu_wavelet(ref, "ref", "fwd", a.copy(), filt)
                    out[key + "_inv"] = helpers.cpu_wavelet(ref, "ref", "inv", a.copy(), filt)
            # multi-level
            a = rng.integers(-300, 301, size=(64, 96)).astype(dtype)
            key = f"{tag}_f{filt}_ml3"
            out[key + "_in"] = a
            out[key + "_fwd"] = helpers.cpu_wavelet(ref, "ref", "fwd", a.copy(), filt, 3)
            out[key + "_inv"] = helpers.cpu_wavelet(ref, "ref", "inv", a.copy(), filt, 3)
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "wavelet.npz"), **out)
    print("wavelet.npz:", len(out), "arrays")


def frame_golden(ref):
    rng = np.random.default_rng(42)
    out = {}
    for idx, (w, h, ext) in enumerate(((20, 20, 4), (8, 8, 2), (5, 3, 3), (64, 48, 32), (9, 2, 8), (1, 1, 2))):
        img = rng.integers(0, 256, size=(h, w)).astype(np.uint8)
        pl = helpers.HostPlane(w, h, ext=ext, upsampled=True, fill=0x55)
        pl.set_image(img)
        helpers.cpu_edgeextend(ref, "ref", pl)
        helpers.cpu_upsample(ref, "ref", pl)
        out[f"up{idx}_img"] = img
        out[f"up{idx}_ext"] = np.array([ext])
        for p in range(4):
            out[f"up{idx}_phase{p}"] = pl.phase(p).copy()
    for idx, (w, h) in enumerate(((10, 10), (39, 39), (11, 7), (2, 2), (1, 5), (135, 99))):
        img = rng.integers(0, 256, size=(h, w)).astype(np.uint8)
        out[f"down{idx}_img"] = img
        out[f"down{idx}_out"] = helpers.cpu_downsample(ref, "ref", img)
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "frame.npz"), **out)
    print("frame.npz:", len(out), "arrays")


OBMC_GOLDEN_CASES = [
    dict(width=64, height=48), dict(width=72, height=40, prec=0), dict(width=72, height=40, prec=1),
    dict(width=72, height=40, prec=3), dict(width=64, height=48, weights=(3, 1, 2)),
    dict(width=64, height=48, weights=(2, 3, 3)), dict(width=64, height=48, num_refs=1),
    dict(width=96, height=64, xbsep=16, ybsep=16, xblen=24, yblen=24),
    dict(width=64, height=48, xbsep=4, ybsep=4, xblen=6, yblen=6, chroma_format=0),
    dict(width=64, height=48, res_is_s32=True), dict(width=66, height=50, span=300, outliers=0.05),
]


def motion_golden(ref):
    """Inputs are regenerated from the seed by the test (ObmcCase is deterministic given the
    oracle's upsampler, itself pinned by frame.npz); only the reference OUTPUTS are stored."""
    oracle = helpers.load_oracle()
    out = {}
    for idx, kw in enumerate(OBMC_GOLDEN_CASES):
        for add in (1, 0):
            if not add and kw.get("res_is_s32"):
                continue
            case = helpers.ObmcCase(oracle, rng=np.random.default_rng(1000 + idx), **kw)
            res = helpers.ref_obmc(ref, case, add)
            res2 = helpers.ref_obmc(ref, case, add, use_ref_renderer=True)
            for k in range(3):
                for q, name in enumerate(("acc", "resid", "out")):
                    if q == 2 and not add:
                        continue
                    out[f"c{idx}_add{add}_k{k}_{name}"] = res[k][q]
                    if name != "acc" and not kw.get("res_is_s32"):   # golden per-pixel renderer agrees (it has no s32 path)
                        assert np.array_equal(res[k][q], res2[k][q]), (kw, add, k, name)
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "motion.npz"), **out)
    print("motion.npz:", len(out), "arrays")


HBM_GOLDEN_CASES = [(128, 96, 3, 0, 0, (5, 3)), (176, 144, 4, 0, 0, (5, 3)), (100, 70, 2, 0, 0, (-7, 4)),
                    (128, 96, 3, 1, 1, (2, -6)), (352, 288, 4, 0, 0, (9, -5))]


def hbm_golden(ref):
    out = {}
    for idx, (w, h, lv, uc, ri, pan) in enumerate(HBM_GOLDEN_CASES):
        s, r = helpers.panning_pair(w, h, np.random.default_rng(2000 + idx), pan)
        fields, pyr = helpers.ref_hbm(ref, s, r, w, h, levels=lv, use_chroma=uc, ref_index=ri)
        out[f"h{idx}_fields"] = fields
        for l in range(lv):
            out[f"h{idx}_pyr{l}"] = pyr[l][0]
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "hbm.npz"), **out)
    print("hbm.npz:", len(out), "arrays")


# (width, height, levels, ref_index, pan, nohint distance, hint distance): sizes the block grid covers exactly
ROUGH_GOLDEN_CASES = [(256, 128, 2, 0, (5, 3), 12, 4), (512, 256, 3, 1, (-7, 2), 12, 4), (640, 384, 4, 0, (11, -6), 12, 4),
                      (256, 192, 3, 0, (4, -4), 7, 2), (384, 256, 2, 0, (20, 9), 20, 6)]


def rough_golden(ref):
    """schro_rough_me_heirarchical_scan of the compiled reference"""
    out = {}
    for idx, (w, h, lv, ri, pan, dn, dh) in enumerate(ROUGH_GOLDEN_CASES):
        s, r = helpers.panning_pair(w, h, np.random.default_rng(3000 + idx), pan)
        out[f"r{idx}_fields"] = helpers.ref_rough(ref, s, r, w, h, levels=lv, ref_index=ri, nohint_distance=dn,
                                                  hint_distance=dh)
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "rough.npz"), **out)
    print("rough.npz:", len(out), "arrays")


SHIFT_MD5_CASES = [(1, 70, 46, 3), (1, 128, 64, 1), (2, 70, 46, 5), (2, 200, 120, 2), (0, 200, 120, 0), (0, 64, 32, 0), (1, 66, 30, 7)]


def shift_md5_inputs(idx, depth, w, h):
    rng = np.random.default_rng(7000 + idx)
    return helpers.random_planes(rng, depth, w, h, True)


def shift_md5_golden(ref):
    """schro_frame_shift_left / _right and schro_frame_md5 of the compiled reference (4:2:0 frames)"""
    import ctypes
    out = {}
    P, I = ctypes.c_void_p * 3, ctypes.c_int * 3
    for idx, (depth, w, h, shift) in enumerate(SHIFT_MD5_CASES):
        planes = shift_md5_inputs(idx, depth, w, h)
        state = (ctypes.c_uint32 * 4)()
        ref.ref_frame_md5(P(*[a.ctypes.data for a in planes]), I(*[a.strides[0] for a in planes]), depth, w, h, state)
        out[f"m{idx}_md5"] = np.array(list(state), np.uint32)
        if depth == 0:
            continue
        for right in (0, 1):
            if not right and depth != 1:
                continue
            q = [a.copy() for a in planes]
            ref.ref_frame_shift(P(*[a.ctypes.data for a in q]), I(*[a.strides[0] for a in q]), depth, w, h, shift, right)
            for c in range(3):
                out[f"s{idx}_{right}_{c}"] = q[c]
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "shift_md5.npz"), **out)
    print("shift_md5.npz:", len(out), "arrays")


LOWDELAY_GOLDEN_CASES = [(64, 32, 2, 4, 2, 97, 2, 0, 0.0), (96, 64, 3, 3, 4, 640, 3, 0, 0.3), (128, 64, 3, 8, 4, 40, 1, 1, 0.0),
                         (96, 96, 2, 6, 6, 33, 1, 1, 0.5), (480, 288, 4, 15, 9, 75, 2, 0, 0.0), (192, 128, 4, 5, 3, 301, 1, 0, 0.2)]


def lowdelay_golden(ref):
    """schro_decoder_decode_lowdelay_transform_data of the compiled reference on slices packed by
    tests/helpers.lowdelay_encode (the bytes are stored: the packer draws random padding)."""
    from tests.test_oracle_lowdelay import quantised_planes
    out = {}
    for idx, (w, h, depth, nh, nv, num, denom, is_s32, trunc) in enumerate(LOWDELAY_GOLDEN_CASES):
        rng = np.random.default_rng(8000 + idx)
        qm = [int(v) for v in rng.integers(0, 8, size=1 + 3 * depth)]
        aligned = ((w // 2) >> depth) % nh == 0 and ((h // 2) >> depth) % nv == 0
        data, _ = helpers.lowdelay_encode(quantised_planes(rng, w, h, 300), depth, nh, nv, num, denom, rng, truncate=trunc,
                                          fast_lengths=aligned and not is_s32)
        planes = helpers.cpu_lowdelay(ref, "ref", data, w, h, depth, nh, nv, num, denom, qm, is_s32, 0)
        out[f"l{idx}_data"] = np.frombuffer(data, np.uint8)
        out[f"l{idx}_qm"] = np.array(qm, np.int32)
        for c in range(3):
            out[f"l{idx}_out{c}"] = planes[c]
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "lowdelay.npz"), **out)
    print("lowdelay.npz:", len(out), "arrays")


def glue_golden(ref):
    """schro_frame_convert / schro_frame_add / schro_frame_subtract of the compiled reference."""
    rng = np.random.default_rng(20261019)
    out = {}
    for sd, dd in ((0, 1), (0, 2), (1, 0), (2, 0), (1, 2), (2, 1)):
        for (sw, sh), (dw, dh) in (((34, 22), (34, 22)), ((34, 22), (28, 18)), ((34, 22), (40, 24))):
            src = helpers.random_planes(rng, sd, sw, sh, True)
            if sd == 2:                       # keep the saturation thresholds of the narrowing in the fixture
                src[0].flat[:8] = [-129, -128, 127, 128, 32639, 32640, 65407, 65408]
            want = helpers.ref_convert(ref, src, sd, sw, sh, dd, dw, dh)
            key = f"conv_{sd}{dd}_{sw}x{sh}_{dw}x{dh}"
            for c in range(3):
                out[f"{key}_in{c}"] = src[c]
                out[f"{key}_out{c}"] = want[c]
    for sd in (0, 1):
        for sub in (0, 1):
            for (sw, sh), (dw, dh) in (((34, 22), (34, 22)), ((30, 20), (34, 22)), ((34, 22), (30, 20))):
                dst = helpers.random_planes(rng, 1, dw, dh, True)
                src = helpers.random_planes(rng, sd, sw, sh, True)
                want = helpers.cpu_add(ref, "ref", dst, dw, dh, src, sd, sw, sh, sub)
                key = f"add_{sd}_{sub}_{sw}x{sh}_{dw}x{dh}"
                for c in range(3):
                    out[f"{key}_dst{c}"] = dst[c]
                    out[f"{key}_src{c}"] = src[c]
                    out[f"{key}_out{c}"] = want[c]
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "glue.npz"), **out)
    print("glue.npz:", len(out), "arrays")


def dequant_golden(ref):
    """orc_dequantise_s16_ip_2d / _s32_ip_2d over the reference's subband / codeblock geometry."""
    from tests.test_oracle_dequant import CASES, make_case
    rng = np.random.default_rng(20261020)
    tables = helpers.ref_quant_tables(ref)
    out = {"table_quant": tables[0], "table_offset_1_2": tables[1], "table_offset_3_8": tables[2]}
    i = 0
    for dtype in (np.int16, np.int32):
        for (w, h, depth, hcb, vcb) in CASES:
            for full in (False, True):
                a, quant = make_case(rng, dtype, w, h, depth, hcb, vcb, tables, full)
                out[f"c{i}_in"], out[f"c{i}_quant"] = a, quant
                out[f"c{i}_depth"], out[f"c{i}_hcb"], out[f"c{i}_vcb"] = np.int32(depth), np.array(hcb), np.array(vcb)
                out[f"c{i}_out"] = helpers.cpu_dequantise(ref, "ref", a, depth, hcb, vcb, quant)
                if dtype == np.int16:
                    # the widening variant (sb2_dequantise_widen): the reference's s32 program on the
                    # sign-extended s16 input
                    out[f"c{i}_wide"] = helpers.cpu_dequantise(ref, "ref", a.astype(np.int32), depth, hcb, vcb, quant)
                i += 1
    out["ncases"] = np.int32(i)
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "dequant.npz"), **out)
    print("dequant.npz:", i, "cases")


def metric_scan_golden(ref):
    """schro_metric_scan_setup / _do_scan / _get_min and schro_metric_fast_block of the compiled reference"""
    from tests import test_oracle_metric_scan as t
    t.REF = ref
    src, rf = t.pictures()
    out = {f"src{k}": src[k] for k in range(3)}
    out.update({f"ref{k}": rf[k] for k in range(3)})
    out["queries"] = np.array(t.QUERIES, np.int32)
    for i, q in enumerate(t.QUERIES):
        o, m, c = t.ref_scan(src, rf, q)
        n = t.used(o)
        out[f"q{i}_out"], out[f"q{i}_metrics"], out[f"q{i}_chroma"] = o, m[:n], c[:n]
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "metric_scan.npz"), **out)
    print("metric_scan.npz:", len(t.QUERIES), "queries")


# (width, height, mv_precision, num_refs, lambda, seed)
SPLIT2_GOLDEN_CASES = [(96, 80, 2, 2, 0.1, 1), (96, 80, 3, 2, 1.0, 2), (128, 64, 1, 2, 0.3, 3), (100, 70, 2, 1, 0.25, 4),
                       (96, 80, 0, 2, 0.05, 5)]


def split2_inputs(oracle, idx, case):
    """pictures from the seed; the sub-pel fields come from the fixture (they are outputs of the oracle's search)"""
    w, h, prec, nrefs, lam, seed = case
    return helpers.split2_case(oracle, w, h, np.random.default_rng(9000 + seed), prec, nrefs, lam=lam)


def split2_golden(ref):
    """schro_do_split2 + schro_motion_copy_to of the compiled reference for every superblock (oracle/ref_me_static.c)"""
    ref_me = helpers.load_ref_me()
    oracle = helpers.load_oracle()
    out = {}
    for idx, case in enumerate(SPLIT2_GOLDEN_CASES):
        w, h, prec, nrefs, lam, seed = case
        src, refs, fields = split2_inputs(oracle, idx, case)
        motion, sb_error, sb_entropy = helpers.ref_split2(ref_me, src, refs, fields, w, h, 8, 8, prec, lam)
        for r, f in enumerate(fields):
            out[f"s{idx}_field{r}"] = f
        out[f"s{idx}_motion"], out[f"s{idx}_sb_error"], out[f"s{idx}_sb_entropy"] = motion, sb_error, sb_entropy
    np.savez_compressed(os.path.join(helpers.GOLDEN_DIR, "split2.npz"), **out)
    print("split2.npz:", len(out), "arrays")


def main():
    ref = helpers.load_ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libschro_ref.so missing: run `make ref` where /root/reference exists")
    wavelet_golden(ref)
    for name in ("frame_golden", "motion_golden", "hbm_golden", "rough_golden", "shift_md5_golden", "lowdelay_golden", "glue_golden", "dequant_golden", "metric_scan_golden", "split2_golden"):
        fn = globals().get(name)
        if fn:
            fn(ref)


if __name__ == "__main__":
    main()
