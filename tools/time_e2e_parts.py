#!/usr/bin/env python
"""e2e throughput of the decode half, the motion-estimation half and both, with N host threads
(development aid: shows which half stops scaling with threads)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bench import HBM_LEVELS
from schroedinger_b200 import lib
nth = int(sys.argv[1]) if len(sys.argv) > 1 else 1
spec = bench.workload_spec("picture_core_2160p")
spec["batch"] = max(2, 2 * nth)
torch.cuda.set_device(0)


class Parts(bench.HostFrames):
    part = "all"

    def picture(self, t, i):
        lib, th = self.lib, self.th[t]
        if self.part in ("all", "decode", "decode_nocopy"):
            if self.part != "decode_nocopy":
                lib.schro_frame_to_gpu(th["coef"], self.coef_host[i])
            lib.schro_frame_inverse_iwt_transform(th["coef"], ctypes.byref(self.params))
            view, cf = th["resid"], th["coef"].contents
            view.regions[0] = cf.regions[0]
            for c in range(3):
                view.components[c].data = cf.components[c].data
            lib.schro_motion_render(th["motion"], th["acc"], ctypes.byref(view), 1, th["out"])
            lib.schro_frame_mc_edgeextend(th["out"])
            th["out"].contents.upsample_done = 0
            lib.schro_upsampled_frame_upsample(th["out"])
            if self.part != "decode_nocopy":
                lib.schro_gpuframe_to_cpu(self.out_host[i], th["out"])
        if self.part in ("all", "me", "me_nocopy"):
            if self.part != "me_nocopy":
                lib.schro_frame_to_gpu(th["src_pyr"][0], self.src_host[i])
            self._build_pyramid(th["src_pyr"])
            arr = self.compat.FrameP * (HBM_LEVELS + 1)
            hbm = lib.schro_hbm_new_from_frames(ctypes.byref(self.params), 0, HBM_LEVELS, 0,
                                                arr(*th["src_pyr"]), arr(*self.ref_pyr))
            lib.schro_hbm_scan(hbm)
            lib.schro_hierarchical_bm_scan_hint(hbm, 0, 3)
            lib.schro_hbm_unref(hbm)
        if self.part == "copy":
            lib.schro_frame_to_gpu(th["coef"], self.coef_host[i])
            lib.schro_frame_to_gpu(th["src_pyr"][0], self.src_host[i])
            lib.schro_gpuframe_to_cpu(self.out_host[i], th["out"])


hf = Parts(spec, lib, nth)
for part in ("all", "decode", "decode_nocopy", "me", "me_nocopy", "copy"):
    hf.part = part
    for _ in range(2):
        hf.step()
    N = 6
    t0 = time.perf_counter()
    for _ in range(N):
        hf.step()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print(f"{nth:3d} threads  {part:14s} {N * spec['batch'] / wall:8.1f} pictures/s  ({wall / (N * spec['batch']) * 1e3:.2f} ms/picture)")
hf.close()
