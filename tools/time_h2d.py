#!/usr/bin/env python
"""Host-to-device rate of schro_frame_to_gpu for the frame kinds the e2e leg uploads (development aid)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from schroedinger_b200 import lib, compat
torch.cuda.set_device(0)
lib.schro_b200_set_device(0)
pinned, cuda = compat.pinned_domain(), compat.cuda_domain()
A = compat.frame_new_and_alloc
W, H, IH = 3840, 2160, 2176


def rate(name, host, devf, nbytes, to_gpu=True, reps=20):
    f = lib.schro_frame_to_gpu if to_gpu else lib.schro_gpuframe_to_cpu
    for _ in range(3):
        f(devf, host) if to_gpu else f(host, devf)
    t0 = time.perf_counter()
    for _ in range(reps):
        f(devf, host) if to_gpu else f(host, devf)
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:58s} {dt * 1e3:7.3f} ms  {nbytes / dt / 1e9:6.1f} GB/s")


cases = [("s16 4:2:0 3840x2176, ext 0 -> ext 0 (plane copies)", compat.FORMAT_S16_420, W, IH, 0, 0, 2),
         ("u8 4:2:0 3840x2160, ext 0 -> ext 32 (row copies)", compat.FORMAT_U8_420, W, H, 0, 32, 1),
         ("u8 4:2:0 3840x2160, ext 0 -> ext 0 (plane copies)", compat.FORMAT_U8_420, W, H, 0, 0, 1),
         ("s16 4:2:0 3840x2176, ext 0 -> ext 8 (row copies)", compat.FORMAT_S16_420, W, IH, 0, 8, 2)]
for name, fmt, w, h, he, de, bpp in cases:
    hf = A(pinned, fmt, w, h, he, 0) if he else A(pinned, fmt, w, h)
    df = A(cuda, fmt, w, h, de, 0)
    nb = int(w * h * 1.5 * bpp)
    rate("H2D " + name, hf, df, nb, True)
    rate("D2H " + name, hf, df, nb, False)
for mb in (12, 25, 100):
    a = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    b = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        b.copy_(a, non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(f"torch pinned -> device {mb} MiB: {dt * 1e3:7.3f} ms {(mb << 20) / dt / 1e9:6.1f} GB/s")
