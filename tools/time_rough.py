"""Device-resident timing of the rough motion search (CUDA events): the full-resolution full search
(the throughput configuration: every 8x8 block of a picture, +-12) and the reference's own chain
(nohint at the coarsest level + hint levels) at 1080p / 2160p."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from schroedinger_b200 import device as dev
from tests import helpers


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    rng = np.random.default_rng(1)
    for (w, h, levels, count) in ((1920, 1080, 4, 32), (3840, 2160, 4, 16)):
        ps, pr = dev.Pyramid(w, h, count, levels), dev.Pyramid(w, h, count, levels)
        for p in range(count):
            s, r = helpers.panning_pair(w, h, rng, (int(rng.integers(-9, 10)), int(rng.integers(-9, 10))))
            for c in range(3):
                ps.slabs[0].upload(p, c, s[c])
                pr.slabs[0].upload(p, c, r[c])
        ps.build()
        pr.build()
        nbx, nby = helpers.hbm_block_counts(w, h, 8, 8)
        prm = dev.HbmParams(8, 8, nbx, nby, 0, 0, 1, 1)
        n = nbx * nby
        out = torch.empty(count * n * 20, dtype=torch.uint8, device="cuda")
        for d in (4, 12, 20):
            ms = timed(lambda: dev.rough_scan_nohint(prm, ps.slabs[0], pr.slabs[0], 0, d, out))
            pos = (2 * d + 1) ** 2
            sads = (w // 8) * (h // 8) * count * pos * 64 / (ms * 1e-3)
            print(f"{w}x{h} x{count}: full search level 0, +-{d}: {ms:.3f} ms  {count / ms * 1e3:.0f} pictures/s  "
                  f"{sads / 1e12:.2f} T byte-SADs/s  ({sads / 4 / 148 / 1.9e9:.1f} VABSDIFF4 lanes/clk/SM at 1.9 GHz)")
        fields = dev.rough_scan(prm, ps, pr)
        ms = timed(lambda: dev.rough_scan(prm, ps, pr, fields=fields))
        print(f"{w}x{h} x{count}: schro_rough_me_heirarchical_scan ({levels} levels): {ms:.3f} ms  {count / ms * 1e3:.0f} pictures/s")


if __name__ == "__main__":
    main()
