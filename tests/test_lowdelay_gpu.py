"""GPU parity: the low-delay slice decoder (sb2_lowdelay_decode) against the oracle and against golden
outputs of the compiled reference (tests/golden/lowdelay.npz), bit-exact: well-formed slices, truncated
and overflowing slices, luma lengths that run into the next slice, arbitrary bytes, values that wrap."""
import os

import numpy as np
import pytest
import torch

from tests import helpers
from tests.golden import make_golden as mg

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()
GOLD = np.load(os.path.join(helpers.GOLDEN_DIR, "lowdelay.npz"))
DQ = np.load(os.path.join(helpers.GOLDEN_DIR, "dequant.npz"))
TABLES = (DQ["table_quant"], DQ["table_offset_1_2"], DQ["table_offset_3_8"])


@pytest.fixture(params=["staged", "unstaged"], autouse=True)
def slice_source(request):
    """Every test runs twice: slices staged in shared memory (default) and read from global memory."""
    from schroedinger_b200 import lib
    lib.sb2_lowdelay_force_unstaged(1 if request.param == "unstaged" else 0)
    yield request.param
    lib.sb2_lowdelay_force_unstaged(0)


def gpu_lowdelay(datas, w, h, depth, nh, nv, num, denom, qm, is_s32):
    """datas: list of per-picture slice buffers (bytes), decoded as one batch."""
    from schroedinger_b200 import device as dev
    count = len(datas)
    nbytes = len(datas[0])
    pitch = (nbytes + 255) // 256 * 256
    buf = np.full(count * pitch, 0xA5, np.uint8)
    for p, d in enumerate(datas):
        buf[p * pitch:p * pitch + nbytes] = np.frombuffer(d, np.uint8)
    slices = torch.from_numpy(buf).cuda()
    coeffs = dev.PictureSlab(dev.FrameLayout.yuv420("s32" if is_s32 else "s16", w, h), count)
    coeffs.buf.fill_(0x5a)
    dev.lowdelay_decode(slices, nbytes, coeffs, depth, nh, nv, num, denom, qm, TABLES[0], TABLES[1], picture_pitch=pitch)
    torch.cuda.synchronize()
    return [[coeffs.download(p, c) for c in range(3)] for p in range(count)]


def oracle_decode(data, w, h, depth, nh, nv, num, denom, qm, is_s32):
    aligned = ((w // 2) >> depth) % nh == 0 and ((h // 2) >> depth) % nv == 0
    return helpers.cpu_lowdelay(ORACLE, "oracle", data, w, h, depth, nh, nv, num, denom, qm, is_s32,
                                1 if (aligned and not is_s32) else 0, TABLES)


def test_lowdelay_golden(cuda):
    for idx, case in enumerate(mg.LOWDELAY_GOLDEN_CASES):
        w, h, depth, nh, nv, num, denom, is_s32 = case[:8]
        data, qm = GOLD[f"l{idx}_data"].tobytes(), [int(v) for v in GOLD[f"l{idx}_qm"]]
        got = gpu_lowdelay([data], w, h, depth, nh, nv, num, denom, qm, is_s32)[0]
        for c in range(3):
            assert np.array_equal(got[c], GOLD[f"l{idx}_out{c}"]), (idx, c)


@pytest.mark.parametrize("case", [
    (64, 32, 2, 4, 2, 97, 2, 0), (96, 64, 3, 3, 4, 640, 3, 0), (128, 64, 3, 8, 4, 40, 1, 0), (128, 64, 3, 8, 4, 40, 1, 1),
    (96, 96, 2, 6, 6, 33, 1, 1), (480, 288, 4, 15, 9, 75, 2, 0), (192, 128, 4, 5, 3, 301, 1, 0)])
def test_lowdelay_vs_oracle(cuda, case):
    w, h, depth, nh, nv, num, denom, is_s32 = case
    rng = np.random.default_rng(sum(case))
    from tests.test_oracle_lowdelay import quantised_planes
    aligned = ((w // 2) >> depth) % nh == 0 and ((h // 2) >> depth) % nv == 0
    fast = aligned and not is_s32
    datas, qms = [], None
    qm = [int(v) for v in rng.integers(0, 8, size=1 + 3 * depth)]
    for (truncate, amp, scale) in ((0.0, 200, 1), (0.5, 200, 1), (0.0, 6000, 4), (0.0, 200, 0.34)):
        q = quantised_planes(rng, w, h, amp)
        n2 = max(4 * denom, int(num * scale))
        if n2 != num:
            continue
        datas.append(helpers.lowdelay_encode(q, depth, nh, nv, num, denom, rng, truncate=truncate, fast_lengths=fast)[0])
    # arbitrary bytes: luma lengths run into the following slices; the last slices are kept inside the buffer
    sizes = helpers.lowdelay_slice_sizes(num, denom, nh, nv)
    raw = rng.integers(0, 256, size=sum(sizes), dtype=np.uint8)
    pos = 0
    for sz in sizes:
        bits = np.unpackbits(raw[pos:pos + sz])
        lb = helpers.ilog2up(8 * ((num // denom) if fast else sz))
        ylen = int("".join(map(str, bits[7:7 + lb])), 2)
        left = 8 * (sum(sizes) - pos) - 7 - lb
        if ylen > left:
            ylen = int(rng.integers(0, left + 1))
            bits[7:7 + lb] = [(ylen >> (lb - 1 - i)) & 1 for i in range(lb)]
        raw[pos:pos + sz] = np.packbits(bits)
        pos += sz
    datas.append(bytes(raw))
    got = gpu_lowdelay(datas, w, h, depth, nh, nv, num, denom, qm, is_s32)
    for p, d in enumerate(datas):
        want = oracle_decode(d, w, h, depth, nh, nv, num, denom, qm, is_s32)
        for c in range(3):
            assert np.array_equal(got[p][c], want[c]), (case, p, c)


def test_lowdelay_1080p_then_inverse_transform(cuda):
    """BASELINE configs[1]: 1080p low-delay intra pictures, DD 9/7 depth 4: slices -> coefficients -> picture."""
    from schroedinger_b200 import device as dev
    from tests.test_oracle_lowdelay import quantised_planes
    w, h, depth, nh, nv, num, denom = 1920, 1088, 4, 60, 34, 190, 1
    rng = np.random.default_rng(12)
    qm = [0, 2, 2, 4, 2, 2, 4, 4, 4, 6, 6, 6, 8]
    q = quantised_planes(rng, w, h, 60)
    data = helpers.lowdelay_encode(q, depth, nh, nv, num, denom, rng, fast_lengths=True)[0]
    want = oracle_decode(data, w, h, depth, nh, nv, num, denom, qm, 0)
    got = gpu_lowdelay([data, data], w, h, depth, nh, nv, num, denom, qm, 0)
    for p in range(2):
        for c in range(3):
            assert np.array_equal(got[p][c], want[c]), (p, c)
    # and on through the inverse wavelet
    layout = dev.FrameLayout.yuv420("s16", w, h)
    a, b = dev.PictureSlab(layout, 1), dev.PictureSlab(layout, 1)
    for c in range(3):
        a.upload(0, c, want[c])
    dev.iwt_inverse(a, b, 0, depth)
    for c in range(3):
        assert np.array_equal(b.download(0, c), helpers.cpu_wavelet(ORACLE, "oracle", "inv", want[c].copy(), 0, depth)), c


SHIM = os.path.join(helpers.ROOT, "oracle", "_ref", "libcompat_shim.so")


@pytest.mark.parametrize("via_shim", [False, True])
@pytest.mark.parametrize("domain_kind", ["malloc", "cuda"])
def test_lowdelay_drop_in(cuda, via_shim, domain_kind):
    """schro_b200_decode_lowdelay_transform_data (and the reference-side schro_decoder_decode_lowdelay_transform_data
    on a SchroPicture) -> schro_frame_inverse_iwt_transform, on a malloc'd and on a CUDA-domain transform frame; the
    library's own quantiser tables (the Dirac specification's formulas) against the reference's."""
    import ctypes
    from schroedinger_b200 import compat, lib
    from tests.test_oracle_lowdelay import quantised_planes
    if via_shim and not os.path.exists(SHIM):
        pytest.skip("oracle/_ref/libcompat_shim.so was not built")
    w, h, depth, nh, nv, num, denom = 480, 288, 4, 15, 9, 150, 1
    rng = np.random.default_rng(77)
    qm = [0, 1, 1, 2, 1, 1, 2, 2, 2, 3, 3, 3, 4]
    data = helpers.lowdelay_encode(quantised_planes(rng, w, h, 100), depth, nh, nv, num, denom, rng, fast_lengths=True)[0]
    want = oracle_decode(data, w, h, depth, nh, nv, num, denom, qm, 0)
    params = compat.make_params(w, h, wavelet_filter_index=0, transform_depth=depth, iwt_luma_width=w, iwt_luma_height=h)
    params.is_lowdelay = 1
    params.n_horiz_slices, params.n_vert_slices = nh, nv
    params.slice_bytes_num, params.slice_bytes_denom = num, denom
    params.iwt_chroma_width, params.iwt_chroma_height = w // 2, h // 2
    for i, v in enumerate(qm):
        params.quant_matrix[i] = v
    domain = compat.cuda_domain() if domain_kind == "cuda" else None
    f = compat.frame_new_and_alloc(domain, compat.FORMAT_S16_420, w, h, 0, 0)
    buf = ctypes.create_string_buffer(data, len(data))
    if via_shim:
        from schroedinger_b200._lib import LIB_PATH
        ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        shim = ctypes.CDLL(SHIM)
        shim.compat_shim_lowdelay.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        shim.compat_shim_lowdelay(ctypes.byref(params), buf, len(data), f)
    else:
        lib.schro_b200_decode_lowdelay_transform_data(ctypes.byref(params), buf, len(data), f)
    host = f
    if domain_kind == "cuda":
        host = compat.frame_new_and_alloc(None, compat.FORMAT_S16_420, w, h, 0, 0)
        lib.schro_gpuframe_to_cpu(host, f)
    for c in range(3):
        assert np.array_equal(np.array(compat.frame_plane(host, c)), want[c]), (c, "coefficients")
    lib.schro_frame_inverse_iwt_transform(f, ctypes.byref(params))
    if domain_kind == "cuda":
        lib.schro_gpuframe_to_cpu(host, f)
    for c in range(3):
        pic = helpers.cpu_wavelet(ORACLE, "oracle", "inv", want[c].copy(), 0, depth)
        assert np.array_equal(np.array(compat.frame_plane(host, c)), pic), (c, "picture")
    if host is not f:
        lib.schro_frame_unref(host)
    lib.schro_frame_unref(f)


def test_lowdelay_batch_drop_in(cuda):
    """schro_b200_decode_lowdelay_pictures: n pictures per call = the per-picture chain, for a shape the fused
    inverse + convert covers and for one it does not."""
    import ctypes
    from schroedinger_b200 import compat, lib
    from tests.test_oracle_lowdelay import quantised_planes
    from tests.test_iwt_convert_gpu import want_pictures
    for (w, h, depth, nh, nv, num, pw, ph, shift) in ((480, 288, 4, 15, 9, 150, 480, 270, 0), (104, 72, 2, 13, 9, 40, 100, 70, 1)):
        rng = np.random.default_rng(w)
        qm = [int(v) for v in rng.integers(0, 4, size=1 + 3 * depth)]
        params = compat.make_params(pw, ph, wavelet_filter_index=0, transform_depth=depth, iwt_luma_width=w, iwt_luma_height=h)
        params.is_lowdelay = 1
        params.n_horiz_slices, params.n_vert_slices = nh, nv
        params.slice_bytes_num, params.slice_bytes_denom = num, 1
        params.iwt_chroma_width, params.iwt_chroma_height = w // 2, h // 2
        for i, v in enumerate(qm):
            params.quant_matrix[i] = v
        n = 5
        aligned = ((w // 2) >> depth) % nh == 0 and ((h // 2) >> depth) % nv == 0
        datas = [helpers.lowdelay_encode(quantised_planes(rng, w, h, 60), depth, nh, nv, num, 1, rng, fast_lengths=aligned)[0]
                 for _ in range(n)]
        bufs = [ctypes.create_string_buffer(d, len(d)) for d in datas]
        outs = [compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, pw, ph, 0, 0) for _ in range(n)]
        lib.schro_b200_decode_lowdelay_pictures(ctypes.byref(params), n, (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in bufs]),
                                                len(datas[0]), (compat.FrameP * n)(*outs), 0, shift)
        for i in range(n):
            coeffs = oracle_decode(datas[i], w, h, depth, nh, nv, num, 1, qm, 0)
            want = want_pictures(coeffs, 0, depth, shift, pw, ph)
            for c in range(3):
                assert np.array_equal(np.array(compat.frame_plane(outs[i], c)), want[c]), ((w, h), i, c)
        for f in outs:
            lib.schro_frame_unref(f)
