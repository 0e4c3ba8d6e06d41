"""compat/schro_hbm_new.c: the reference-side half of the one constructor whose argument the library
does not know (SchroEncoderFrame, schrohierbm.c:25-64).

CPU: compiled against the reference's own headers (where /root/reference exists) it defines
schro_hbm_new and leaves exactly schro_hbm_new_from_frames to the library.
GPU: through oracle/_ref/libcompat_shim.so (the shim + a fixture that builds the reference's
SchroEncoderFrame structures, built here, travels with the snapshot) a schro_hbm_new (frame, 0) call
produces the same motion fields as the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests import helpers

REF = "/root/reference"
OBJ = os.path.join(helpers.ROOT, "oracle", "_ref", "obj", "compat_schro_hbm_new.o")
SHIM = os.path.join(helpers.ROOT, "oracle", "_ref", "libcompat_shim.so")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference headers not present on this box")
def test_shim_compiles_against_reference_headers():
    subprocess.check_call(["bash", os.path.join(helpers.ROOT, "oracle", "build_ref.sh")],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    syms = subprocess.check_output(["nm", OBJ], text=True).split("\n")
    defined = {l.split()[-1] for l in syms if " T " in l}
    undefined = {l.split()[-1] for l in syms if l.strip().startswith("U ")}
    assert defined == {"schro_hbm_new"}
    assert "schro_hbm_new_from_frames" in undefined
    # nothing else of the library or of the reference is needed (the rest is libc + the assert's logger)
    assert {u for u in undefined if u.startswith("schro_")} == {"schro_hbm_new_from_frames", "schro_debug_log"}


def test_library_exports_what_the_shim_needs():
    out = subprocess.check_output(["nm", "-D", os.path.join(helpers.ROOT, "schroedinger_b200", "libschro_b200.so")],
                                  text=True)
    assert " T schro_hbm_new_from_frames" in out
    assert " T schro_hbm_new\n" not in out          # the shim owns that name


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SHIM), reason="oracle/_ref/libcompat_shim.so was not built")
def test_schro_hbm_new_through_the_shim(cuda):
    from schroedinger_b200 import compat, lib
    from tests.test_host_api_gpu import _new_u8_frame
    oracle = helpers.load_oracle()
    from schroedinger_b200._lib import LIB_PATH
    ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)                # the same handle, promoted: schro_hbm_new_from_frames for the shim
    shim = ctypes.CDLL(SHIM)
    shim.compat_shim_hbm_new.restype = ctypes.c_void_p
    shim.compat_shim_free.argtypes = [ctypes.c_void_p]
    w, h, levels = 320, 192, 3
    s, r = helpers.panning_pair(w, h, np.random.default_rng(21), (-3, 5))
    want, _, _ = helpers.oracle_hbm(oracle, s, r, w, h, levels=levels)
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
    fixture = ctypes.c_void_p()
    hbm_addr = shim.compat_shim_hbm_new(ctypes.byref(params), levels, 0, arr(*fs), arr(*fr), ctypes.byref(fixture))
    assert hbm_addr
    hbm = ctypes.cast(hbm_addr, ctypes.POINTER(compat.SchroHierBm))
    assert hbm.contents.hierarchy_levels == levels and hbm.contents.ref == 0
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
    shim.compat_shim_free(fixture)


# ---- compat/schro_rough_me_new.c (schroroughmotion.c:21-33) -----------------------------------------
ROUGH_OBJ = os.path.join(helpers.ROOT, "oracle", "_ref", "obj", "compat_schro_rough_me_new.o")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference headers not present on this box")
def test_rough_shim_compiles_against_reference_headers():
    subprocess.check_call(["bash", os.path.join(helpers.ROOT, "oracle", "build_ref.sh")],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    syms = subprocess.check_output(["nm", ROUGH_OBJ], text=True).split("\n")
    defined = {l.split()[-1] for l in syms if " T " in l}
    undefined = {l.split()[-1] for l in syms if l.strip().startswith("U ")}
    assert defined == {"schro_rough_me_new"}
    assert {u for u in undefined if u.startswith("schro_")} == {"schro_rough_me_new_from_frames", "schro_debug_log"}
    out = subprocess.check_output(["nm", "-D", os.path.join(helpers.ROOT, "schroedinger_b200", "libschro_b200.so")],
                                  text=True)
    assert " T schro_rough_me_new_from_frames" in out and " T schro_rough_me_new\n" not in out
    for name in ("schro_rough_me_free", "schro_rough_me_heirarchical_scan", "schro_rough_me_heirarchical_scan_nohint",
                 "schro_rough_me_heirarchical_scan_hint"):
        assert f" T {name}\n" in out, name


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SHIM), reason="oracle/_ref/libcompat_shim.so was not built")
def test_schro_rough_me_new_through_the_shim(cuda):
    from schroedinger_b200 import compat, lib
    from tests.test_host_api_gpu import _new_u8_frame
    oracle = helpers.load_oracle()
    from schroedinger_b200._lib import LIB_PATH
    ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    shim = ctypes.CDLL(SHIM)
    shim.compat_shim_rough_me_new.restype = ctypes.c_void_p
    shim.compat_shim_free.argtypes = [ctypes.c_void_p]
    w, h, levels = 320, 192, 3
    s, r = helpers.panning_pair(w, h, np.random.default_rng(22), (5, -3))
    want, _, _ = helpers.oracle_rough(oracle, s, r, w, h, levels=levels, ref_index=1)
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
    fixture = ctypes.c_void_p()
    addr = shim.compat_shim_rough_me_new(ctypes.byref(params), levels, 1, arr(*fs), arr(*fr), ctypes.byref(fixture))
    assert addr
    rme = ctypes.cast(addr, ctypes.POINTER(compat.SchroRoughME))
    assert rme.contents.encoder_frame and rme.contents.ref_frame
    lib.schro_rough_me_heirarchical_scan(rme)
    n = params.x_num_blocks * params.y_num_blocks
    for l in range(levels, 0, -1):
        mf = rme.contents.motion_fields[l]
        got = np.ctypeslib.as_array(ctypes.cast(mf.contents.motion_vectors, ctypes.POINTER(ctypes.c_uint8)),
                                    shape=(n * 20,)).view(helpers.MV_DTYPE)
        for f in ("flags", "metric", "chroma_metric", "v"):
            assert np.array_equal(got[f], want[l][f]), (l, f)
    lib.schro_rough_me_free(rme)
    shim.compat_shim_free(fixture)
