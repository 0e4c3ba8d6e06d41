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
