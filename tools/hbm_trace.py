import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from schroedinger_b200 import device as dev, lib
spec = bench.workload_spec("picture_core_2160p"); spec["batch"] = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bench.CONTENT = sys.argv[2] if len(sys.argv) > 2 else "natural"
st = bench.Stages(spec, torch, dev)
for _ in range(3): st.step()
torch.cuda.synchronize()
buf = np.zeros(512 * 8, np.int64)
lib.sb2_hbm_trace_read(buf.ctypes.data_as(ctypes.c_void_p), 512 * 8)
t = buf.reshape(512, 8)[50:450]
d = np.diff(t[:, :7], axis=1)
names = ["A static cands", "poll+dedup", "rank (nbr SADs)", "seed+sync", "scan", "reduce+sync+sel"]
print("batch", spec["batch"], bench.CONTENT, "per-block cycles (row 100, level 0), median / mean:")
for k, n in enumerate(names): print(f"  {n:18s} {np.median(d[:,k]):8.0f} {d[:,k].mean():8.0f}")
tot = np.diff(t[:, 0])
print("  block-to-block     ", np.median(tot), tot.mean())

