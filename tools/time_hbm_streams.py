#!/usr/bin/env python
"""K independent single-picture hierarchical block matches on K streams, launched from one host thread:
does the GPU overlap them?  (development aid)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from schroedinger_b200 import device as dev
W, H = 3840, 2160
nbx, nby = bench.block_counts(W, H)
rng = np.random.default_rng(1)
base = bench.textured_frame(W, H, rng)
srcf = bench.textured_frame(W, H, rng, pan=(5, 3))
KMAX = 16
jobs = []
for k in range(KMAX):
    sp, rp = dev.Pyramid(W, H, 1, bench.HBM_LEVELS, 8), dev.Pyramid(W, H, 1, bench.HBM_LEVELS, 8)
    for pyr, fr in ((sp, srcf), (rp, base)):
        for c in range(3):
            pyr.slabs[0].upload(0, c, fr[c])
        pyr.build()
    fields = [torch.empty(nbx * nby * 20, dtype=torch.uint8, device="cuda") for _ in range(bench.HBM_LEVELS + 1)]
    jobs.append(dict(sp=sp, rp=rp, fields=fields, ws=dev.Workspace(), st=torch.cuda.Stream()))
prm = dev.HbmParams(8, 8, nbx, nby, 0, 0, 1, 1)
torch.cuda.synchronize()
for K in (1, 2, 4, 8, 16):
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for j in jobs[:K]:
            with torch.cuda.stream(j["st"]):
                dev.hbm_scan(prm, j["sp"], j["rp"], 3, j["fields"], j["ws"], stream=j["st"])
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    print(f"K={K:2d}: launch {1e3 * (t1 - t0):6.2f} ms, total {1e3 * (t2 - t0):6.2f} ms -> {K / (t2 - t0):7.1f} pictures/s")
