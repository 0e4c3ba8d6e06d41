"""GPU: sb2_dequantise against the oracle and the golden vectors of the compiled reference,
bit-exact."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.test_oracle_dequant import CASES, make_case

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()
GOLD = np.load(os.path.join(helpers.GOLDEN_DIR, "dequant.npz"))
TABLES = (GOLD["table_quant"], GOLD["table_offset_1_2"], GOLD["table_offset_3_8"])


def gpu_dequant(planes_per_pic, depth, hcb, vcb, quant_per_pic, dtype):
    """planes_per_pic: list (pictures) of lists (components) of arrays; quant likewise."""
    import torch
    from schroedinger_b200 import device as dev
    name = "s32" if dtype == np.int32 else "s16"
    sizes = [(p.shape[1], p.shape[0]) for p in planes_per_pic[0]]
    slab = dev.PictureSlab(dev.FrameLayout(name, sizes), len(planes_per_pic))
    for i, planes in enumerate(planes_per_pic):
        for c, a in enumerate(planes):
            slab.upload(i, c, a)
    q = np.concatenate([np.concatenate([qq.reshape(-1) for qq in qs]) for qs in quant_per_pic]).astype(np.int32)
    dev.dequantise(slab, depth, hcb, vcb, torch.from_numpy(q).cuda())
    return [[slab.download(i, c) for c in range(len(planes))] for i, planes in enumerate(planes_per_pic)]


def test_dequantise_golden(cuda):
    for i in range(int(GOLD["ncases"])):
        depth, hcb, vcb = int(GOLD[f"c{i}_depth"]), GOLD[f"c{i}_hcb"].tolist(), GOLD[f"c{i}_vcb"].tolist()
        a = GOLD[f"c{i}_in"]
        got = gpu_dequant([[a]], depth, hcb, vcb, [[GOLD[f"c{i}_quant"]]], a.dtype)
        assert np.array_equal(got[0][0], GOLD[f"c{i}_out"]), i


@pytest.mark.parametrize("dtype", [np.int16, np.int32])
def test_dequantise_batched_420(cuda, dtype):
    """three pictures x three components (4:2:0 sizes) in one launch, each with its own tables"""
    rng = np.random.default_rng(9)
    for (w, h, depth, hcb, vcb) in CASES[1:4]:
        w2, h2 = 2 * w, 2 * h
        pics, quants, wants = [], [], []
        for _ in range(3):
            planes, qs, ws = [], [], []
            for (pw, ph) in ((w2, h2), (w, h), (w, h)):
                a, q = make_case(rng, dtype, pw, ph, depth, hcb, vcb, TABLES, False)
                planes.append(a); qs.append(q)
                ws.append(helpers.cpu_dequantise(ORACLE, "oracle", a, depth, hcb, vcb, q))
            pics.append(planes); quants.append(qs); wants.append(ws)
        got = gpu_dequant(pics, depth, hcb, vcb, quants, dtype)
        for i in range(3):
            for c in range(3):
                assert np.array_equal(got[i][c], wants[i][c]), (dtype, w, h, i, c)


def test_dequantise_2160p_s32(cuda):
    """BASELINE shape: 3840x2176 s32, 5 levels, the codeblock grid a 2160p stream would carry"""
    rng = np.random.default_rng(10)
    depth, hcb, vcb = 5, [1, 1, 2, 4, 8, 12], [1, 1, 2, 3, 6, 8]
    planes, qs, ws = [], [], []
    for (pw, ph) in ((3840, 2176), (1920, 1088), (1920, 1088)):
        a, q = make_case(rng, np.int32, pw, ph, depth, hcb, vcb, TABLES, False)
        planes.append(a); qs.append(q)
        ws.append(helpers.cpu_dequantise(ORACLE, "oracle", a, depth, hcb, vcb, q))
    got = gpu_dequant([planes], depth, hcb, vcb, [qs], np.int32)
    for c in range(3):
        assert np.array_equal(got[0][c], ws[c]), c


@pytest.mark.parametrize("domain_kind", ["malloc", "cuda"])
def test_frame_dequantise_then_inverse_transform(cuda, domain_kind):
    """The decoder's picture tail through the drop-in layer: quantised coefficients ->
    schro_b200_frame_dequantise -> schro_frame_inverse_iwt_transform, on host and CUDA-domain frames."""
    import ctypes
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(77)
    w, h, depth, filt = 352, 288, 4, 1
    params = compat.make_params(w, h, filt, depth)
    hcb, vcb = [1, 1, 2, 3, 4], [1, 1, 2, 2, 3]
    for i in range(depth + 1):
        params.horiz_codeblocks[i], params.vert_codeblocks[i] = hcb[i], vcb[i]
    iw, ih = params.iwt_luma_width, params.iwt_luma_height
    host = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, iw, ih)
    want, pairs = [], []
    for c in range(3):
        v = compat.frame_plane(host, c)
        a, q = make_case(rng, np.int16, v.shape[1], v.shape[0], depth, hcb, vcb, TABLES, False)
        v[...] = a
        pairs.append(q)
        d = helpers.cpu_dequantise(ORACLE, "oracle", a, depth, hcb, vcb, q)
        want.append(helpers.cpu_wavelet(ORACLE, "oracle", "inv", d, filt, depth))
    table = np.ascontiguousarray(np.concatenate([q.reshape(-1) for q in pairs]).astype(np.int32))
    if domain_kind == "cuda":
        dom = compat.cuda_domain()
        work = compat.frame_new_and_alloc(dom, compat.FORMAT_S16_420, iw, ih)
        lib.schro_frame_to_gpu(work, host)
    else:
        work = host
    lib.schro_b200_frame_dequantise(work, ctypes.byref(params), table.ctypes.data_as(ctypes.c_void_p))
    lib.schro_frame_inverse_iwt_transform(work, ctypes.byref(params))
    if work is not host:
        lib.schro_gpuframe_to_cpu(host, work)
    for c in range(3):
        assert np.array_equal(compat.frame_plane(host, c), want[c]), (domain_kind, c)


# ---- widening variant: s16 quantised coefficients -> s32 coefficient frame --------------------
def gpu_dequant_widen(planes_per_pic, depth, hcb, vcb, quant_per_pic):
    import torch
    from schroedinger_b200 import device as dev
    sizes = [(p.shape[1], p.shape[0]) for p in planes_per_pic[0]]
    src = dev.PictureSlab(dev.FrameLayout("s16", sizes), len(planes_per_pic))
    dst = dev.PictureSlab(dev.FrameLayout("s32", sizes), len(planes_per_pic))
    for i, planes in enumerate(planes_per_pic):
        for c, a in enumerate(planes):
            src.upload(i, c, a)
    q = np.concatenate([np.concatenate([qq.reshape(-1) for qq in qs]) for qs in quant_per_pic]).astype(np.int32)
    dev.dequantise_widen(src, dst, depth, hcb, vcb, torch.from_numpy(q).cuda())
    return ([[dst.download(i, c) for c in range(len(planes))] for i, planes in enumerate(planes_per_pic)],
            [[src.download(i, c) for c in range(len(planes))] for i, planes in enumerate(planes_per_pic)])


def test_dequantise_widen_golden(cuda):
    """the compiled reference's s32 program on the sign-extended s16 golden inputs (full-range values
    included); the s16 source is left untouched"""
    n = 0
    for i in range(int(GOLD["ncases"])):
        if f"c{i}_wide" not in GOLD.files:
            continue
        depth, hcb, vcb = int(GOLD[f"c{i}_depth"]), GOLD[f"c{i}_hcb"].tolist(), GOLD[f"c{i}_vcb"].tolist()
        a = GOLD[f"c{i}_in"]
        got, src_after = gpu_dequant_widen([[a]], depth, hcb, vcb, [[GOLD[f"c{i}_quant"]]])
        assert got[0][0].dtype == np.int32 and np.array_equal(got[0][0], GOLD[f"c{i}_wide"]), i
        assert np.array_equal(src_after[0][0], a), i
        n += 1
    assert n >= 10


def test_dequantise_widen_2160p(cuda):
    """BASELINE shape, two pictures x three components, against the oracle's s32 program"""
    rng = np.random.default_rng(12)
    depth, hcb, vcb = 5, [1, 1, 2, 4, 8, 12], [1, 1, 2, 3, 6, 8]
    pics, quants, wants = [], [], []
    for _ in range(2):
        planes, qs, ws = [], [], []
        for (pw, ph) in ((3840, 2176), (1920, 1088), (1920, 1088)):
            a, q = make_case(rng, np.int16, pw, ph, depth, hcb, vcb, TABLES, False)
            planes.append(a); qs.append(q)
            ws.append(helpers.cpu_dequantise(ORACLE, "oracle", a.astype(np.int32), depth, hcb, vcb, q))
        pics.append(planes); quants.append(qs); wants.append(ws)
    got, _ = gpu_dequant_widen(pics, depth, hcb, vcb, quants)
    for i in range(2):
        for c in range(3):
            assert np.array_equal(got[i][c], wants[i][c]), (i, c)


@pytest.mark.parametrize("domain_kind", ["malloc", "pinned_to_cuda"])
def test_frame_dequantise_widen_then_inverse_transform(cuda, domain_kind):
    """What the e2e leg does per picture: s16 quantised coefficients (host) -> device s32 coefficient
    frame -> inverse transform, through the drop-in layer."""
    import ctypes
    from schroedinger_b200 import compat, lib
    rng = np.random.default_rng(78)
    w, h, depth, filt = 352, 288, 4, 6
    params = compat.make_params(w, h, filt, depth)
    hcb, vcb = [1, 1, 2, 3, 4], [1, 1, 2, 2, 3]
    for i in range(depth + 1):
        params.horiz_codeblocks[i], params.vert_codeblocks[i] = hcb[i], vcb[i]
    iw, ih = params.iwt_luma_width, params.iwt_luma_height
    src_dom = compat.pinned_domain() if domain_kind != "malloc" else None
    dst_dom = compat.cuda_domain() if domain_kind != "malloc" else None
    src = compat.frame_new_and_alloc(src_dom, compat.FORMAT_S16_420, iw, ih)
    dst = compat.frame_new_and_alloc(dst_dom, compat.FORMAT_S32_420, iw, ih)
    out = compat.frame_new_and_alloc(None, compat.FORMAT_S32_420, iw, ih)
    want, pairs = [], []
    for c in range(3):
        v = compat.frame_plane(src, c)
        a, q = make_case(rng, np.int16, v.shape[1], v.shape[0], depth, hcb, vcb, TABLES, False)
        v[...] = a
        pairs.append(q)
        d = helpers.cpu_dequantise(ORACLE, "oracle", a.astype(np.int32), depth, hcb, vcb, q)
        want.append(helpers.cpu_wavelet(ORACLE, "oracle", "inv", d, filt, depth))
    table = np.ascontiguousarray(np.concatenate([q.reshape(-1) for q in pairs]).astype(np.int32))
    lib.schro_b200_frame_dequantise_widen(dst, src, ctypes.byref(params), table.ctypes.data_as(ctypes.c_void_p))
    lib.schro_frame_inverse_iwt_transform(dst, ctypes.byref(params))
    if dst_dom is not None:
        lib.schro_gpuframe_to_cpu(out, dst)
    else:
        out = dst
    for c in range(3):
        assert np.array_equal(compat.frame_plane(out, c), want[c]), (domain_kind, c)
