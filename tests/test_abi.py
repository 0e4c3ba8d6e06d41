"""CPU: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls without a GPU)."""
import ctypes
import glob
import os
import re

from tests import helpers


def declared_symbols():
    names = set()
    for path in glob.glob(os.path.join(helpers.ROOT, "include", "*.h")):
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b((?:sb2|schro)_[a-z0-9_]+)\s*\(", text):
            names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import schroedinger_b200 as s
    syms = declared_symbols()
    assert len(syms) >= 6
    missing = [n for n in syms if not hasattr(s.lib, n)]
    assert not missing, missing


def test_version_and_error_string():
    import schroedinger_b200 as s
    assert s.lib.sb2_version() >= 1
    assert isinstance(s.last_error(), str)


def test_argument_errors_do_not_need_a_gpu():
    import schroedinger_b200 as s
    slab = s.Slab()
    rc = s.lib.sb2_iwt_forward(ctypes.byref(slab), ctypes.byref(slab), 0, 1, 1, None, 0, None)
    assert rc != 0
    assert "slab" in s.last_error()


def test_e2e_driver_loads_and_exports_its_entry_points():
    """bench.py's pthread driver (bench_native/e2e_driver.c) links against the drop-in library;
    it must load without a GPU (nothing runs until a job is started)."""
    path = os.path.join(helpers.ROOT, "bench_native", "libsb2_e2e_driver.so")
    assert os.path.exists(path), "run `make` (or __graft_entry__.build())"
    drv = ctypes.CDLL(path)
    for name in ("sb2_e2e_start", "sb2_e2e_run", "sb2_e2e_step", "sb2_e2e_times", "sb2_e2e_stop"):
        assert hasattr(drv, name), name


def _slab3(s, w, h, count=1):
    """a three-component 4:2:0 slab description with a null base: enough for the argument checks, which run
    before anything touches the device"""
    slab = s.Slab()
    slab.base = 0x1000
    slab.count = count
    slab.ncomp = 3
    for k, (cw, ch) in enumerate(((w, h), ((w + 1) // 2, (h + 1) // 2), ((w + 1) // 2, (h + 1) // 2))):
        slab.width[k], slab.height[k], slab.stride[k] = cw, ch, (cw + 64 + 15) & ~15
    return slab


def test_split2_argument_errors_do_not_need_a_gpu():
    """sb2_split2_decide validates geometry before it launches anything (include/schro_b200.h)."""
    import schroedinger_b200 as s
    from schroedinger_b200._lib import Split2Params
    w, h = 64, 48
    orig, up = _slab3(s, w, h), _slab3(s, w, h)
    dummy = ctypes.c_void_p(0x1000)

    def call(p, o=orig, u0=up, u1=up, f1=dummy, ws=dummy, ws_bytes=1 << 20):
        return s.lib.sb2_split2_decide(ctypes.byref(p), ctypes.byref(o), ctypes.byref(u0), ctypes.byref(u1) if u1 else None, 32,
                                       dummy, f1, 64, dummy, 64, dummy, dummy, ws, ws_bytes, None)

    good = lambda: Split2Params(8, 8, 8, 8, 2, 2, 1, 1, 32, 0.1)
    p = good(); p.num_refs = 3
    assert call(p) != 0 and "num_refs" in s.last_error()
    p = good()
    assert call(p, u1=None) != 0 and "num_refs" in s.last_error()          # two references announced, one given
    p = good(); p.x_num_blocks = 6
    assert call(p) != 0 and "bad parameters" in s.last_error()             # superblocks are 4 x 4 blocks
    p = good(); p.mv_precision = 4
    assert call(p) != 0 and "bad parameters" in s.last_error()
    p = good(); p.y_num_blocks = 1028
    assert call(p) != 0 and "1024" in s.last_error()
    p = good()
    assert call(p, u0=_slab3(s, w + 2, h)) != 0 and "differ in size" in s.last_error()
    p = good(); p.chroma_h_shift = 0
    assert call(p) != 0 and "chroma" in s.last_error()
    p = good()
    assert call(p, ws_bytes=16) != 0 and "workspace" in s.last_error()
    assert s.lib.sb2_split2_workspace_bytes(8, 8, 3) == 8 * 8 * 3 * 32


def test_subpel_argument_errors_do_not_need_a_gpu():
    import schroedinger_b200 as s
    from schroedinger_b200._lib import SubpelParams
    w, h = 64, 48
    orig, up = _slab3(s, w, h), _slab3(s, w, h)
    dummy = ctypes.c_void_p(0x1000)

    def call(p, u=up, ext=32, ws_bytes=1 << 20):
        return s.lib.sb2_subpel_refine(ctypes.byref(p), ctypes.byref(orig), ctypes.byref(u), ext, dummy, 64, dummy, ws_bytes, None)

    good = lambda: SubpelParams(8, 8, 8, 8, 2, 0, 32, 0.1)
    p = good(); p.ref_index = 2
    assert call(p) != 0 and "bad parameters" in s.last_error()
    p = good(); p.y_num_blocks = 2000
    assert call(p) != 0 and "1024" in s.last_error()
    p = good()
    assert call(p, u=_slab3(s, w, h + 2)) != 0 and "differ in size" in s.last_error()
    p = good()
    assert call(p, ext=4) != 0 and "border" in s.last_error()              # probes would leave the reference's border
    p = good()
    assert call(p, ws_bytes=16) != 0 and "workspace" in s.last_error()
    assert s.lib.sb2_subpel_workspace_bytes(8, 8, 2) == 8 * 8 * 2 * 48
