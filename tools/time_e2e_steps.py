#!/usr/bin/env python
"""Per-step wall time of bench.py's e2e arm with N host threads (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from schroedinger_b200 import lib
nthreads = int(sys.argv[1]) if len(sys.argv) > 1 else 4
spec = bench.workload_spec("picture_core_2160p")
torch.cuda.set_device(0)
hf = bench.HostFrames(spec, lib, nthreads)
ts = []
for i in range(25):
    t = time.perf_counter(); hf.step(); ts.append(time.perf_counter() - t)
print(nthreads, "threads; step ms:", " ".join(f"{x*1e3:.0f}" for x in ts))
print("fps (last 15 steps):", spec["batch"] * 15 / sum(ts[-15:]))
