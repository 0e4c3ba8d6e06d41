#!/usr/bin/env python
"""Do PCIe copies and the motion-estimation kernels overlap?  Half of the host threads only copy, the
other half only run pyramid + block matching on resident frames (development aid)."""
import sys, os, time, ctypes, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bench import HBM_LEVELS
from schroedinger_b200 import lib
nth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spec = bench.workload_spec("picture_core_2160p")
spec["batch"] = 2 * nth
torch.cuda.set_device(0)
hf = bench.HostFrames(spec, lib, nth)


def copy_only(t, i):
    th = hf.th[t]
    lib.schro_frame_to_gpu(th["coef"], hf.coef_host[i])
    lib.schro_frame_to_gpu(th["src_pyr"][0], hf.src_host[i])
    lib.schro_gpuframe_to_cpu(hf.out_host[i], th["out"])


def me_only(t, i):
    th = hf.th[t]
    hf._build_pyramid(th["src_pyr"])
    arr = hf.compat.FrameP * (HBM_LEVELS + 1)
    hbm = lib.schro_hbm_new_from_frames(ctypes.byref(hf.params), 0, HBM_LEVELS, 0, arr(*th["src_pyr"]), arr(*hf.ref_pyr))
    lib.schro_hbm_scan(hbm)
    lib.schro_hierarchical_bm_scan_hint(hbm, 0, 3)
    lib.schro_hbm_unref(hbm)


def run(assign, seconds=3.0):
    counts = [0] * nth
    stop = time.perf_counter() + seconds

    def w(t):
        fn = assign(t)
        if fn is None:
            return
        while time.perf_counter() < stop:
            fn(t, t)
            counts[t] += 1
        lib.schro_b200_thread_release()
    ths = [threading.Thread(target=w, args=(t,)) for t in range(nth)]
    t0 = time.perf_counter()
    for x in ths: x.start()
    for x in ths: x.join()
    dt = time.perf_counter() - t0
    return [c / dt for c in counts]


half = nth // 2
for name, assign in (("copy threads only", lambda t: copy_only if t < half else None),
                     ("ME threads only", lambda t: me_only if t >= half else None),
                     ("both at once", lambda t: copy_only if t < half else me_only)):
    r = run(assign)
    print(f"{name:20s} copies {sum(r[:half]):7.1f}/s   block matches {sum(r[half:]):7.1f}/s")
