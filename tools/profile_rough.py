"""One launch of the full-resolution full search (1080p, 32 pictures, +-12) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from schroedinger_b200 import device as dev
from tests import helpers

w, h, count = 1920, 1080, 32
rng = np.random.default_rng(1)
ps, pr = dev.Pyramid(w, h, count, 0), dev.Pyramid(w, h, count, 0)
s, r = helpers.panning_pair(w, h, rng, (5, 3))
for p in range(count):
    for c in range(3):
        ps.slabs[0].upload(p, c, s[c])
        pr.slabs[0].upload(p, c, r[c])
ps.build(); pr.build()
nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
out = torch.empty(count * nbx * nby * 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    dev.rough_scan_nohint(dev.HbmParams(8, 8, nbx, nby, 0, 0, 1, 1), ps.slabs[0], pr.slabs[0], 0, 12, out)
torch.cuda.synchronize()
