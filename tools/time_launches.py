"""Per-launch CUDA-event times of the pyramid build (downsample chain) at the bench size."""
import ctypes
import sys
import torch
sys.path.insert(0, ".")
from schroedinger_b200 import device as dev, lib

W, H, COUNT, LEVELS = 3840, 2160, int(sys.argv[1]) if len(sys.argv) > 1 else 32, 4
pyr = dev.Pyramid(W, H, COUNT, LEVELS)
pyr.slabs[0].buf.random_(0, 256)
for _ in range(3):
    pyr.build()
torch.cuda.synchronize()
lib.sb2_profile_reset()
lib.sb2_profile_enable(1)
for _ in range(5):
    pyr.build()
torch.cuda.synchronize()
lib.sb2_profile_enable(0)
buf = ctypes.create_string_buffer(64)
ms, by = ctypes.c_float(), ctypes.c_double()
n = lib.sb2_profile_count()
per = n // 5
for i in range(n - per, n):
    lib.sb2_profile_get(i, buf, 64, ctypes.byref(ms), ctypes.byref(by))
    print(f"{buf.value.decode():24s} {ms.value*1000:8.1f} us  {by.value/1e6:8.1f} MB  {by.value/ms.value/1e6:7.1f} GB/s")
