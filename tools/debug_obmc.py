"""Debug aid: run one golden OBMC case on a forced kernel and print where it differs from the oracle."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests import helpers
from tests.golden import make_golden as mg
from tests import test_obmc_gpu as T
from schroedinger_b200 import lib

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 3
kern = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kw = mg.OBMC_GOLDEN_CASES[idx]
ORACLE = helpers.load_oracle()
case = helpers.ObmcCase(ORACLE, rng=np.random.default_rng(1000 + idx), **kw)
lib.sb2_obmc_force_kernel(kern)
got = T.gpu_obmc(case, 1)[0]
want = helpers.oracle_obmc(ORACLE, case, 1)
for k in range(3):
    d = got[k][0].astype(int) - want[k][0].astype(int)
    ys, xs = np.nonzero(d)
    print("comp", k, "mismatches", len(ys), "of", d.size)
    for y, x in list(zip(ys, xs))[:40]:
        print("  y", y, "x", x, "got", got[k][0][y, x], "want", want[k][0][y, x], "diff", d[y, x])
