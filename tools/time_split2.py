"""Device-resident timing of the split-2 pass of the mode decision (CUDA events, per kernel through sb2_profile)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from schroedinger_b200 import device as dev, lib
from tests import helpers


def main():
    rng = np.random.default_rng(1)
    for (w, h, count) in ((1920, 1080, 32), (3840, 2160, 16)):
        nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
        orig = dev.PictureSlab(dev.FrameLayout.yuv420("u8", w, h, 32), count)
        ups = [dev.PictureSlab(dev.FrameLayout.yuv420("u8", w, h, 32, True), count) for _ in range(2)]
        s, r0 = helpers.panning_pair(w, h, rng, (5, 3))
        _, r1 = helpers.panning_pair(w, h, rng, (-4, 2))
        for p in range(count):
            for c in range(3):
                orig.upload(p, c, s[c])
                ups[0].upload(p, c, r0[c])
                ups[1].upload(p, c, r1[c])
        for u in ups:
            dev.edgeextend_upsample(u)
        flds = []
        for r, pan in enumerate(((5, 3), (-4, 2))):
            f = np.zeros(count * nbx * nby, helpers.MV_DTYPE)
            f["flags"] = r + 1
            f["v"][:, r] = -4 * pan[0] + rng.integers(-2, 3, size=len(f))
            f["v"][:, 2 + r] = -4 * pan[1] + rng.integers(-2, 3, size=len(f))
            f["metric"] = rng.integers(100, 2000, size=len(f))
            flds.append(torch.from_numpy(f.view(np.uint8).copy()).cuda())
        for _ in range(2):
            dev.split2_decide(orig, ups, flds, 8, 8, nbx, nby, 2, 0.1)
        torch.cuda.synchronize()
        lib.sb2_profile_reset(); lib.sb2_profile_enable(1)
        reps = 5
        for _ in range(reps):
            m, e, n = dev.split2_decide(orig, ups, flds, 8, 8, nbx, nby, 2, 0.1)
        torch.cuda.synchronize()
        lib.sb2_profile_enable(0)
        buf = ctypes.create_string_buffer(64); ms = ctypes.c_float(); by = ctypes.c_double()
        tot = {}
        for i in range(lib.sb2_profile_count()):
            lib.sb2_profile_get(i, buf, 64, ctypes.byref(ms), ctypes.byref(by))
            t = tot.setdefault(buf.value.decode(), [0.0, 0.0, 0]); t[0] += ms.value; t[1] += by.value; t[2] += 1
        lib.sb2_profile_reset()
        total = sum(v[0] for v in tot.values()) / reps
        modes = m.cpu().numpy().view(helpers.MV_DTYPE)["flags"] & 3
        print(f"{w}x{h} x{count}, two references, quarter-pel: {total:.3f} ms = {count / total * 1e3:.0f} pictures/s; "
              f"modes dc/ref0/ref1/both = {[int((modes == k).sum()) for k in range(4)]}")
        for k, v in sorted(tot.items()):
            print(f"   {k}: {v[0] / v[2]:.3f} ms per launch, {v[1] / v[2] / (v[0] / v[2]) * 1e-6:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
