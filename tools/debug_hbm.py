"""Development aid: run one block-matching case through the wavefront kernel, the generic kernel and
the oracle, print where they differ (level, block, fields)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests import helpers
from tests.test_hbm_gpu import gpu_hbm
from schroedinger_b200 import lib

ORACLE = helpers.load_oracle()


def run(w, h, lv, pan, seed, noise=3):
    s, r = helpers.panning_pair(w, h, np.random.default_rng(seed), pan, noise=noise)
    want, _, _ = helpers.oracle_hbm(ORACLE, s, r, w, h, levels=lv)
    nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
    for kind in ("wave", "generic"):
        lib.sb2_hbm_force_generic(1 if kind == "generic" else 0)
        got, _ = gpu_hbm([(s, r)], w, h, lv)
        got = got[0]
        for l in range(lv, -1, -1):
            bad = np.nonzero((got[l]["metric"] != want[l]["metric"]) | (got[l]["v"] != want[l]["v"]).any(axis=1)
                             | (got[l]["flags"] != want[l]["flags"]))[0]
            print(f"{kind} {w}x{h} level {l}: {len(bad)} of {nbx * nby} entries differ")
            for b in bad[:6]:
                print(f"   block ({b % nbx},{b // nbx}) got v={got[l]['v'][b]} m={got[l]['metric'][b]} "
                      f"want v={want[l]['v'][b]} m={want[l]['metric'][b]}")
    lib.sb2_hbm_force_generic(0)


if __name__ == "__main__":
    if hasattr(lib, "sb2_hbm_wave_debug") and len(sys.argv) > 3:
        lib.sb2_hbm_wave_debug(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
    run(128, 96, 3, (5, 3), 2000)
