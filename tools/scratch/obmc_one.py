import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tests import helpers
from tests.test_obmc_gpu import gpu_obmc
from schroedinger_b200 import lib
ORACLE = helpers.load_oracle()
lib.sb2_obmc_force_kernel(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
case = helpers.ObmcCase(ORACLE, rng=np.random.default_rng(77), width=352, height=288, span=200, outliers=0.05)
want = helpers.oracle_obmc(ORACLE, case, 1)
got = gpu_obmc(case, 1, count=1)
torch.cuda.synchronize()
for k in range(3):
    for q in range(3):
        print(k, q, np.array_equal(got[0][k][q], want[k][q]))
