#!/usr/bin/env python
"""Wall time of each drop-in API call of one e2e picture, single host thread (development aid)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from schroedinger_b200 import lib
spec = bench.workload_spec("picture_core_2160p")
nth = int(sys.argv[1]) if len(sys.argv) > 1 else 1
spec["batch"] = max(2, 2 * nth)
torch.cuda.set_device(0)
hf = bench.HostFrames(spec, lib, nth)
for _ in range(3):
    hf.step()
acc = {}
real = {}
names = ["schro_frame_to_gpu", "schro_frame_inverse_iwt_transform", "schro_motion_render", "schro_frame_mc_edgeextend",
         "schro_upsampled_frame_upsample", "schro_gpuframe_to_cpu", "schro_frame_downsample", "schro_hbm_new_from_frames",
         "schro_hbm_scan", "schro_hierarchical_bm_scan_hint", "schro_hbm_unref"]


class Timed:
    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, n):
        f = getattr(self._lib, n)
        if n not in names:
            return f

        def g(*a):
            t = time.perf_counter()
            r = f(*a)
            acc[n] = acc.get(n, 0.0) + time.perf_counter() - t
            acc[n + "#"] = acc.get(n + "#", 0) + 1
            return r
        return g


hf.lib = Timed(lib)
N = 10
t0 = time.perf_counter()
for _ in range(N):
    hf.step()
tot = time.perf_counter() - t0
npic = N * spec["batch"]
print(f"{nth} thread(s): {tot / npic * 1e3:.2f} ms per picture wall = {npic / tot:.0f} fps; per-call times below are per picture, summed over threads")
for n in names:
    if n in acc:
        print(f"  {n:36s} {acc[n] / npic * 1e3:7.3f} ms/picture in {acc[n + '#'] / npic:.0f} calls")
