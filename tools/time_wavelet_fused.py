#!/usr/bin/env python
"""NS-1 measurement: the inverse transform with its last two levels fused into one launch
(wavelet_inv_fused2_kernel) next to the one-launch-per-level path, same inputs, CUDA events, plus the
per-kernel times of both (sb2_profile)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from schroedinger_b200 import device as dev, lib


def kernel_times(fn, reps):
    lib.sb2_profile_reset(); lib.sb2_profile_enable(1)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    lib.sb2_profile_enable(0)
    buf = ctypes.create_string_buffer(64); ms = ctypes.c_float(); by = ctypes.c_double()
    tot = {}
    for i in range(lib.sb2_profile_count()):
        lib.sb2_profile_get(i, buf, 64, ctypes.byref(ms), ctypes.byref(by))
        t = tot.setdefault(buf.value.decode(), [0.0, 0]); t[0] += ms.value; t[1] += 1
    lib.sb2_profile_reset()
    return {k: v[0] / reps for k, v in tot.items()}


def run(name, depth_name, filt, depth, w, h, count, iters=20):
    layout = dev.FrameLayout.yuv420(depth_name, w, h)
    a = dev.PictureSlab(layout, count, zero=False)
    b = dev.PictureSlab(layout, count, zero=False)
    a.buf.random_(0, 255)
    ncoef = sum(cw * ch for cw, ch in layout.comp_sizes)
    alg = 2 * ncoef * layout.bpp * count
    for fused in (0, 1):
        lib.sb2_iwt_enable_fused(fused)
        for _ in range(3):
            dev.iwt_inverse(a, b, filt, depth)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            dev.iwt_inverse(a, b, filt, depth)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"{name:26s} {'fused 1+0' if fused else 'per level':10s} {count:3d} pics {ms:7.3f} ms {count / ms * 1e3:9.0f} pics/s "
              f"{alg / ms / 1e6:7.0f} GB/s algorithmic = {alg / ms / 1e6 / 6545.9 * 100:5.1f}% of 6545.9")
        kt = kernel_times(lambda: dev.iwt_inverse(a, b, filt, depth), 5)
        print("      " + "  ".join(f"{k.replace('wavelet_inv_', '')}={v:.3f}" for k, v in sorted(kt.items())))
    lib.sb2_iwt_enable_fused(0)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "only-d2":
        # under ncu: one configuration, few iterations (every launch is replayed)
        run("Daub 9/7 s32 2160p d2", "s32", 6, 2, 3840, 2176, 32, iters=2)
        sys.exit(0)
    run("Daub 9/7 s32 2160p d5", "s32", 6, 5, 3840, 2176, 32)
    run("Daub 9/7 s32 2160p d2", "s32", 6, 2, 3840, 2176, 32)
    run("DD 9/7 s16 1080p d4", "s16", 0, 4, 1920, 1088, 64)
    run("DD 13/7 s16 1080p d4", "s16", 2, 4, 1920, 1088, 64)
    run("DD 9/7 s32 1080p d4", "s32", 0, 4, 1920, 1088, 64)
