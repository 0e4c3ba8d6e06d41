"""schroedinger_b200 -- B200-native picture core for Dirac / VC-2 (schroedinger drop-in).

The package is a thin ctypes binding over ``libschro_b200.so`` (hand-written sm_100a
CUDA kernels + the C host layer mirroring the reference's C API).  There is NO CPU
fallback: importing works without a GPU (so the ABI can be inspected), but every
compute call requires the CUDA library and a device and fails loudly otherwise.
"""
from ._lib import lib, LIB_PATH, Sb2Error, Slab, check, last_error  # noqa: F401
from . import device  # noqa: F401

__all__ = ["lib", "LIB_PATH", "Sb2Error", "Slab", "check", "last_error", "device"]
