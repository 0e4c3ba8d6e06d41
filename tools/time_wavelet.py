#!/usr/bin/env python
"""Device-resident timing of the wavelet transforms, the convert glue and the dequantiser (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from schroedinger_b200 import device as dev

def run(name, depth_name, filt, depth, w, h, count, inverse, iters=10):
    layout = dev.FrameLayout.yuv420(depth_name, w, h)
    a = dev.PictureSlab(layout, count, zero=False)
    b = dev.PictureSlab(layout, count, zero=False)
    a.buf.random_(0, 255)
    fn = dev.iwt_inverse if inverse else dev.iwt_forward
    for _ in range(3):
        fn(a, b, filt, depth)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn(a, b, filt, depth)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ncoef = sum(cw * ch for cw, ch in layout.comp_sizes)
    alg = 2 * ncoef * layout.bpp * count
    print(f"{name:28s} {count:4d} pics {ms:8.3f} ms  {count/ms*1e3:9.1f} pics/s  "
          f"{alg/ms/1e6:8.1f} GB/s algorithmic ({alg/ms/1e6/6545.9*100:5.1f}% of 6545.9)")

def run_glue():
    """combine / convert glue: bytes moved per second against the HBM peak"""
    for (sd, dd, w, h, count) in (("s16", "u8", 1920, 1080, 64), ("s32", "u8", 3840, 2160, 16), ("u8", "s16", 1920, 1080, 64)):
        a = dev.PictureSlab(dev.FrameLayout.yuv420(sd, w, h), count, zero=False)
        b = dev.PictureSlab(dev.FrameLayout.yuv420(dd, w, h), count, zero=False)
        a.buf.random_(0, 255)
        for _ in range(3):
            dev.frame_convert(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dev.frame_convert(a, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = w * h * 1.5 * (a.layout.bpp + b.layout.bpp) * count
        print(f"convert {sd}->{dd} {w}x{h}       {count:4d} pics {ms:8.3f} ms  {count/ms*1e3:9.1f} pics/s  "
              f"{nbytes/ms/1e6:8.1f} GB/s ({nbytes/ms/1e6/6545.9*100:5.1f}% of 6545.9)")



def run_dequant():
    """dequantisation in place"""
    import numpy as np
    for (name, w, h, count, depth) in (("s32", 3840, 2176, 16, 5), ("s16", 1920, 1088, 64, 4)):
        a = dev.PictureSlab(dev.FrameLayout.yuv420(name, w, h), count, zero=False)
        a.buf.random_(0, 8)
        hcb = [1, 1, 2, 4, 8, 12][:depth + 1]
        vcb = [1, 1, 2, 3, 6, 8][:depth + 1]
        n = hcb[0] * vcb[0] + sum(3 * hcb[l + 1] * vcb[l + 1] for l in range(depth))
        q = torch.tensor([[64, 34]] * (n * 3 * count), dtype=torch.int32, device="cuda").reshape(-1)
        for _ in range(3):
            dev.dequantise(a, depth, hcb, vcb, q)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dev.dequantise(a, depth, hcb, vcb, q)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = w * h * 1.5 * 2 * a.layout.bpp * count
        print(f"dequantise {name} {w}x{h} d{depth}     {count:4d} pics {ms:8.3f} ms  {count/ms*1e3:9.1f} pics/s  "
              f"{nbytes/ms/1e6:8.1f} GB/s ({nbytes/ms/1e6/6545.9*100:5.1f}% of 6545.9)")


if __name__ == "__main__":
    run("LeGall s16 1080p d4 fwd", "s16", 1, 4, 1920, 1088, 64, False)
    run("LeGall s16 1080p d4 inv", "s16", 1, 4, 1920, 1088, 64, True)
    run("DD9/7 s16 1080p d4 inv", "s16", 0, 4, 1920, 1088, 64, True)
    run("DD9/7 s16 1080p d4 fwd", "s16", 0, 4, 1920, 1088, 64, False)
    run("Daub s32 2160p d5 inv", "s32", 6, 5, 3840, 2176, 8, True)
    run("Daub s32 2160p d5 fwd", "s32", 6, 5, 3840, 2176, 8, False)
    run("Daub s32 2160p d1 inv", "s32", 6, 1, 3840, 2176, 8, True)
    run("Fidelity s16 1080p d4 inv", "s16", 5, 4, 1920, 1088, 64, True)
    run("Haar0 s16 1080p d4 inv", "s16", 3, 4, 1920, 1088, 64, True)
    run_glue()
    run_dequant()
