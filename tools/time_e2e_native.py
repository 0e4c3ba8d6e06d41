#!/usr/bin/env python
"""Where the host threads of the native e2e driver spend their time (development aid)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from schroedinger_b200 import lib
nth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spec = bench.workload_spec("picture_core_2160p")
spec["batch"] = 2 * nth
torch.cuda.set_device(0)
hf = bench.HostFrames(spec, lib, nth)
hf.native_start()
for _ in range(3):
    hf.step()
names = ["H2D coefficients", "inverse wavelet", "motion render", "edge-extend + upsample", "D2H picture", "H2D source",
         "pyramid", "hbm_new", "hbm_scan", "scan_hint level 0", "hbm_unref"]
out = (ctypes.c_double * 16)()
hf.drv.sb2_e2e_times(1, None)
N, wall = 10, 0.0
hf.drv.sb2_e2e_step.restype = ctypes.c_double
for _ in range(N):
    wall += hf.drv.sb2_e2e_step()
hf.drv.sb2_e2e_times(0, out)
npic = N * spec["batch"]
print(f"{nth} threads: {npic / wall:.0f} fps, {wall / npic * 1e3:.2f} ms/picture wall, {wall * nth / npic * 1e3:.2f} ms/picture thread time")
for k, n in enumerate(names):
    print(f"  {n:24s} {out[k] / npic * 1e3:7.3f} ms/picture")
hf.close()
