#!/usr/bin/env python
"""OBMC render time for the bench's random vector field (every block its own mode and
sub-pel phase -- worst case for a warp that spans two blocks) against a coherent field (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from schroedinger_b200 import device as dev, lib

orig = bench.make_mv_field


def coherent(nbx, nby, rng):
    mv = orig(nbx, nby, rng)
    mv["flags"] = 3                                     # both references everywhere
    mv["v"] = np.array([13, -22, 7, 5], np.int16)       # one quarter-pel vector pair (all four taps live)
    return mv


for name, fn in (("random field (bench, C4)", orig), ("coherent field", coherent)):
    bench.make_mv_field = fn
    spec = bench.workload_spec("picture_core_2160p"); spec["batch"] = 8
    st = bench.Stages(spec, torch, dev)
    stage = [s for s in st.stages if s["name"] == "obmc_render"][0]
    for _ in range(3): stage["run"]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): stage["run"]()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / 10:.3f} ms per 8 pictures")
    del st
