#!/usr/bin/env python
"""Threaded throughput of subsets of the e2e picture stages (SB2_E2E_STAGES mask of bench_native/e2e_driver.c):
which stage stops scaling with the number of host threads?  (development aid)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mask = int(sys.argv[2]) if len(sys.argv) > 2 else 255
os.environ["SB2_E2E_STAGES"] = str(mask)
import torch
import bench
from schroedinger_b200 import lib
spec = bench.workload_spec("picture_core_2160p")
spec["batch"] = max(32, nth)
torch.cuda.set_device(0)
hf = bench.HostFrames(spec, lib, nth)
hf.native_start()
for _ in range(2):
    hf.step()
import resource
N = 20
ru0 = resource.getrusage(resource.RUSAGE_SELF)
wall = hf.drv.sb2_e2e_run(N)
ru1 = resource.getrusage(resource.RUSAGE_SELF)
cpu = (ru1.ru_utime - ru0.ru_utime) + (ru1.ru_stime - ru0.ru_stime)
npic = N * spec["batch"]
names = {1: "H2D coef", 2: "decode kernels", 4: "D2H picture", 8: "H2D source", 16: "pyramid", 32: "block matching"}
what = "+".join(v for k, v in names.items() if mask & k)
print(f"{nth:2d} threads, stages [{what}]: {npic / wall:7.0f} pictures/s, {wall * nth / npic * 1e3:6.2f} ms per picture per thread, "
      f"host CPU {cpu / npic * 1e3:.2f} ms per picture ({cpu / wall:.1f} cores busy)")
hf.close()
