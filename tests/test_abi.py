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
