"""The low-delay e2e leg of bench.py on its own, for several host-thread counts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from schroedinger_b200 import device as dev, lib
torch.cuda.set_device(0)
for n in [int(a) for a in sys.argv[1:]] or [16, 32, 48, 64]:
    r = bench.lowdelay_rows(torch, dev, lib, 1, lambda v, op: v, lambda: torch.cuda.synchronize(), e2e_threads=n)["lowdelay_1080p"]
    print(n, "threads: device-resident", r["value"], "frames/s, e2e", r["e2e"]["value"], "frames/s", r["e2e_call_ms"], "runs", r["e2e"]["runs"], "batched", r["e2e_batched"]["value"], r["e2e_batched"]["runs"])
