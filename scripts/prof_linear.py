"""Run two masked method='linear' interpolations of one config (for ncu): python scripts/prof_linear.py c1"""
import sys

import torch

sys.path.insert(0, ".")
from ptv_interpolation_b200 import synthetic
from ptv_interpolation_b200.engine import PTVEngine

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
planes = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # only the first `planes` z-planes (large configs)
dev = torch.device("cuda", 0)
eng = PTVEngine(dev)
cfg = synthetic.make_config(name, device=dev)
n = cfg["n"]
mask = cfg["mask"].view(torch.uint8)
ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)
eng.build(cfg["points"], cfg["values"])
az = ax[:planes] if planes else ax
if planes:
    mask = mask[:planes].contiguous()
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = eng.interpolate(ax, ax, az, mask=mask, method="linear")
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} linear {e0.elapsed_time(e1):.2f} ms")
