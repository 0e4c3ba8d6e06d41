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


def main():
    ref = helpers.load_ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libschro_ref.so missing: run `make ref` where /root/reference exists")
    wavelet_golden(ref)
    for name in ("frame_golden", "motion_golden", "hbm_golden"):
        fn = globals().get(name)
        if fn:
            fn(ref)


if __name__ == "__main__":
    main()
