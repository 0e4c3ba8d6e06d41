"""Slice-decode kernel time against the number of pictures per launch (CUDA events via sb2_profile)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from schroedinger_b200 import device as dev, lib
torch.cuda.set_device(0)
w, h, depth, nh, nv, nbytes = 1920, 1088, 4, 60, 34, 190
rng = np.random.default_rng(4242)
templates = [bench._lowdelay_slice(rng, nbytes, 32 * 32, 2 * 16 * 16, int(rng.integers(8, 28))) for _ in range(16)]
pic_bytes = nh * nv * nbytes
pitch = (pic_bytes + 255) // 256 * 256
qm = [0, 2, 2, 4, 2, 2, 4, 4, 4, 6, 6, 6, 8]
tq = [4] * 61; to = [1] * 61
for count in (1, 2, 8, 64):
    host = np.zeros((count, pitch), np.uint8)
    for p in range(count):
        pick = rng.integers(0, len(templates), size=nh * nv)
        host[p, :pic_bytes] = np.concatenate([templates[k] for k in pick])
    slices = torch.from_numpy(host.reshape(-1)).cuda()
    coeffs = dev.PictureSlab(dev.FrameLayout.yuv420("s16", w, h), count)
    for _ in range(3):
        dev.lowdelay_decode(slices, pic_bytes, coeffs, depth, nh, nv, nbytes, 1, qm, tq, to, picture_pitch=pitch)
    torch.cuda.synchronize()
    lib.sb2_profile_reset(); lib.sb2_profile_enable(1)
    for _ in range(5):
        dev.lowdelay_decode(slices, pic_bytes, coeffs, depth, nh, nv, nbytes, 1, qm, tq, to, picture_pitch=pitch)
    torch.cuda.synchronize(); lib.sb2_profile_enable(0)
    prof = bench.collect_profile(lib); lib.sb2_profile_reset()
    print(count, "pictures:", {k: round(v["ms"] / v["launches"], 4) for k, v in prof.items()})
