"""GPU parity: schro_frame_shift_left / _right (sb2_frame_shift) and schro_frame_md5 against the compiled
reference's outputs (tests/golden/shift_md5.npz, made by tests/golden/make_golden.py) and a numpy
restatement of the two Orc programs (schroorc.orc:146-163, 207-218)."""
import ctypes
import os

import numpy as np
import pytest

from tests import helpers
from tests.golden import make_golden as mg

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(helpers.GOLDEN_DIR, "shift_md5.npz"))
FMT = {0: "FORMAT_U8_420", 1: "FORMAT_S16_420", 2: "FORMAT_S32_420"}


def np_shift(a, shift, right):
    if not right:
        return (a.astype(np.int64) << shift).astype(a.dtype)                       # shlw wraps
    rnd = (1 << shift) >> 1
    return ((a.astype(np.int64) + rnd).astype(a.dtype).astype(np.int64) >> shift).astype(a.dtype)   # addw / addl wrap, then shrs


def _frame(compat, fmt_name, w, h, planes, domain=None):
    f = compat.frame_new_and_alloc(domain, getattr(compat, fmt_name), w, h, 0, 0)
    for c in range(3):
        compat.frame_plane(f, c)[...] = planes[c]
    return f


@pytest.mark.parametrize("idx", range(len(mg.SHIFT_MD5_CASES)))
def test_shift_and_md5_match_reference(cuda, idx):
    from schroedinger_b200 import compat, lib
    depth, w, h, shift = mg.SHIFT_MD5_CASES[idx]
    planes = mg.shift_md5_inputs(idx, depth, w, h)
    if not hasattr(compat, FMT[depth]):
        pytest.skip("format constant missing")
    f = _frame(compat, FMT[depth], w, h, planes)
    state = (ctypes.c_uint32 * 4)()
    lib.schro_frame_md5(f, state)
    assert list(state) == list(GOLD[f"m{idx}_md5"]), "md5 of a host frame"
    if depth:
        for right in (0, 1):
            if not right and depth != 1:
                continue
            for c in range(3):
                compat.frame_plane(f, c)[...] = planes[c]
            (lib.schro_frame_shift_right if right else lib.schro_frame_shift_left)(f, shift)
            for c in range(3):
                got = np.array(compat.frame_plane(f, c))
                assert np.array_equal(got, GOLD[f"s{idx}_{right}_{c}"]), (idx, right, c)
                assert np.array_equal(got, np_shift(planes[c], shift, right)), (idx, right, c, "numpy")
    lib.schro_frame_unref(f)


def test_shift_batch_device_resident(cuda):
    """sb2_frame_shift on a slab of pictures, full-range values, every shift."""
    from schroedinger_b200 import device as dev, lib
    from schroedinger_b200._lib import check
    rng = np.random.default_rng(3)
    for name, dtype, depth in (("s16", np.int16, 1), ("s32", np.int32, 2)):
        layout = dev.FrameLayout.yuv420(name, 200, 72)
        slab = dev.PictureSlab(layout, 3)
        info = np.iinfo(dtype)
        for shift in (0, 1, 4, 9):
            for right in (1, 0):
                if not right and depth != 1:
                    continue
                planes = {}
                for p in range(3):
                    for k, (w, h) in enumerate(layout.comp_sizes):
                        planes[p, k] = rng.integers(info.min, info.max + 1, size=(h, w)).astype(dtype)
                        slab.upload(p, k, planes[p, k])
                check(lib.sb2_frame_shift(ctypes.byref(slab.slab), depth, shift, right, None), "sb2_frame_shift")
                for (p, k), a in planes.items():
                    assert np.array_equal(slab.download(p, k), np_shift(a, shift, right)), (name, shift, right, p, k)


def test_md5_of_a_cuda_domain_frame(cuda):
    from schroedinger_b200 import compat, lib
    idx = 4
    depth, w, h, _ = mg.SHIFT_MD5_CASES[idx]
    planes = mg.shift_md5_inputs(idx, depth, w, h)
    host = _frame(compat, FMT[depth], w, h, planes)
    gpu = compat.frame_new_and_alloc(compat.cuda_domain(), getattr(compat, FMT[depth]), w, h, 0, 0)
    lib.schro_frame_to_gpu(gpu, host)
    state = (ctypes.c_uint32 * 4)()
    lib.schro_frame_md5(gpu, state)
    assert list(state) == list(GOLD[f"m{idx}_md5"])
    lib.schro_frame_unref(gpu)
    lib.schro_frame_unref(host)
