#!/usr/bin/env python
"""Per-kernel device time (CUDA events on the launching stream) while the e2e arm runs with N host threads:
shows how much each kernel stretches when pictures of different threads share the GPU (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from schroedinger_b200 import lib
nth = int(sys.argv[1]) if len(sys.argv) > 1 else 1
spec = bench.workload_spec("picture_core_2160p")
spec["batch"] = max(2, 2 * nth)
torch.cuda.set_device(0)
hf = bench.HostFrames(spec, lib, nth)
for _ in range(3):
    hf.step()
lib.sb2_profile_reset()
lib.sb2_profile_enable(1)
N = 4
t0 = time.perf_counter()
for _ in range(N):
    hf.step()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
lib.sb2_profile_enable(0)
prof = bench.collect_profile(lib)
npic = N * spec["batch"]
print(f"{nth} thread(s): {npic / wall:.0f} fps, wall {wall / npic * 1e3:.2f} ms/picture (profiling on)")
tot = 0
for k, r in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:34s} {r['ms'] / npic:8.3f} ms/picture  {r['launches'] / npic:5.1f} launches/picture  {r['ms'] / r['launches']:7.3f} ms/launch")
    tot += r["ms"]
print(f"  sum of kernel durations {tot / npic:.3f} ms/picture -> average kernel concurrency {tot / (wall * 1e3):.2f}")
hf.close()
