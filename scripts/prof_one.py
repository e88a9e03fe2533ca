"""Run a few masked interpolations of one config (for ncu): python scripts/prof_one.py c2 [key=val ...]"""
import sys

import torch

sys.path.insert(0, ".")
from ptv_interpolation_b200 import synthetic
from ptv_interpolation_b200.engine import PTVEngine, set_tuning

name = sys.argv[1]
for a in sys.argv[2:]:
    k, v = a.split("=")
    set_tuning(**{k: float(v)})
dev = torch.device("cuda", 0)
eng = PTVEngine(dev)
cfg = synthetic.make_config(name, device=dev)
n = cfg["n"]
mask = cfg["mask"].view(torch.uint8)
ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)
eng.build(cfg["points"], cfg["values"])
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = eng.interpolate(ax, ax, ax, mask=mask, method=cfg["method"], k=cfg["k"])
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} interp {e0.elapsed_time(e1):.2f} ms")
