#!/usr/bin/env python
"""Hierarchical block matching of B pictures as G groups on G streams: do the under-occupied coarse levels
of one group overlap with the fine levels of another?  (development aid)"""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from schroedinger_b200 import device as dev
from schroedinger_b200._lib import Slab

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W, H = 3840, 2160
nbx, nby = bench.block_counts(W, H)
rng = np.random.default_rng(1)
base = bench.textured_frame(W, H, rng)
srcf = bench.textured_frame(W, H, rng, pan=(5, 3))
sp, rp = dev.Pyramid(W, H, B, bench.HBM_LEVELS, 8), dev.Pyramid(W, H, B, bench.HBM_LEVELS, 8)
for pyr, fr in ((sp, srcf), (rp, base)):
    for c in range(3):
        pyr.slabs[0].upload(0, c, fr[c])
    l0 = pyr.slabs[0]
    one = l0.buf[:l0.layout.pitch]
    for p in range(1, B):
        l0.buf[p * l0.layout.pitch:(p + 1) * l0.layout.pitch].copy_(one)
    pyr.build()
n = nbx * nby
fields = [torch.empty(B * n * 20, dtype=torch.uint8, device="cuda") for _ in range(bench.HBM_LEVELS + 1)]
prm = dev.HbmParams(8, 8, nbx, nby, 0, 0, 1, 1)


class Group:
    """pictures [start, start + count) of a slab"""
    def __init__(self, parent, start, count):
        self.layout, self.count = parent.layout, count
        s = Slab()
        for f, _ in Slab._fields_:
            setattr(s, f, getattr(parent.slab, f))
        s.base = parent.slab.base + start * parent.layout.pitch
        s.count = count
        self.slab = s


class PyrGroup:
    def __init__(self, pyr, start, count):
        self.levels = pyr.levels
        self.slabs = [Group(s, start, count) for s in pyr.slabs]


torch.cuda.synchronize()
for G in [g for g in (1, 2, 4, 8, 16, 32) if g <= B]:
    per = B // G
    jobs = []
    for g in range(G):
        jobs.append(dict(sp=PyrGroup(sp, g * per, per), rp=PyrGroup(rp, g * per, per),
                         fields=[f[g * per * n * 20:(g + 1) * per * n * 20] for f in fields],
                         ws=dev.Workspace(), st=torch.cuda.Stream()))
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cur = torch.cuda.current_stream()
        evs = []
        for j in jobs:                       # fork every stream first: joining inside this loop would chain them
            j["st"].wait_stream(cur)
        for j in jobs:
            dev.hbm_scan(prm, j["sp"], j["rp"], 3, j["fields"], j["ws"], stream=j["st"])
        for j in jobs:
            cur.wait_stream(j["st"])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"B={B} groups={G}: {best:.3f} ms")
