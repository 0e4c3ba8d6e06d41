"""Development aid: per-step cycle breakdown of the wavefront block matcher (clock64 trace of one warp,
level 0, row group 10, picture 0).  Build with `make EXTRA_NVFLAGS=-DSB2_HBM_TRACE` first.
usage: python tools/hbm_trace.py [batch] [content]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from schroedinger_b200 import device as dev, lib
spec = bench.workload_spec("picture_core_2160p"); spec["batch"] = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bench.CONTENT = sys.argv[2] if len(sys.argv) > 2 else "natural"
level = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lib.sb2_hbm_wave_trace_select(level, 10 if level < 3 else 4)
st = bench.Stages(spec, torch, dev)
for _ in range(3): st.step()
torch.cuda.synchronize()
buf = np.zeros(1024 * 8, np.int64)
lib.sb2_hbm_wave_trace_read(buf.ctypes.data_as(ctypes.c_void_p), 1024 * 8)
ncol = {0: 480, 1: 240, 2: 120, 3: 60, 4: 30}[level]
t = buf.reshape(1024, 8)[min(20, ncol // 4):ncol - 2]
d = np.diff(t[:, :7], axis=1)
names = ["neighbours (shfl / poll)", "match + twins", "rank (neighbour SADs)", "winner + window", "src rows + scan", "min-reduce"]
print("batch", spec["batch"], bench.CONTENT, f"per-step cycles (level {level}), median / mean:")
for k, n in enumerate(names): print(f"  {n:26s} {np.median(d[:,k]):8.0f} {d[:,k].mean():8.0f}")
tot = np.diff(t[:, 0])
print("  step-to-step              ", np.median(tot), tot.mean())
