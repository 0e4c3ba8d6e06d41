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
