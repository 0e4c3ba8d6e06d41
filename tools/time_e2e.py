#!/usr/bin/env python
"""Per-call timing of the drop-in API chain used by bench.py's e2e arm (development aid)."""
import sys, os, time, ctypes, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from schroedinger_b200 import lib

spec = bench.workload_spec("picture_core_2160p")
spec["batch"] = 4
torch.cuda.set_device(0)
hf = bench.HostFrames(spec, lib, 1)
T = collections.defaultdict(float)
orig = {}
names = ["schro_frame_to_gpu", "schro_frame_inverse_iwt_transform", "schro_motion_render",
         "schro_frame_mc_edgeextend", "schro_upsampled_frame_upsample", "schro_gpuframe_to_cpu",
         "schro_frame_downsample", "schro_hbm_new_from_frames", "schro_hbm_scan",
         "schro_hierarchical_bm_scan_hint", "schro_hbm_unref"]
class Timed:
    def __init__(self, name, fn): self.name, self.fn = name, fn
    def __call__(self, *a):
        t = time.perf_counter(); r = self.fn(*a); T[self.name] += time.perf_counter() - t; return r
class LibProxy:
    def __getattr__(self, n):
        f = getattr(lib, n)
        return Timed(n, f) if n in names else f
hf.lib = LibProxy()
for _ in range(2): hf.step()
T.clear()
N = 3
t0 = time.perf_counter()
for _ in range(N): hf.step()
tot = time.perf_counter() - t0
pics = N * spec["batch"]
print(f"single thread: {tot/pics*1e3:.2f} ms per picture ({pics/tot:.1f} fps)")
for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
    print(f"  {k:40s} {v/pics*1e3:8.3f} ms/picture")

lib.sb2_profile_reset(); lib.sb2_profile_enable(1)
t0 = time.perf_counter()
for _ in range(N): hf.step()
torch.cuda.synchronize()
lib.sb2_profile_enable(0)
prof = bench.collect_profile(lib)
print("kernel time per picture (CUDA events):")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:40s} {v['ms']/pics:8.3f} ms/picture  ({v['launches']//pics} launches)")
