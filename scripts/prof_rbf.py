"""Local RBF (method='rbf', k=20) on config 2 (256^3, 1M vectors), pore voxels only: timing / ncu target."""
import json, sys, torch
sys.path.insert(0, ".")
from ptv_interpolation_b200 import synthetic
from ptv_interpolation_b200.engine import PTVEngine
dev = torch.device("cuda", 0)
eng = PTVEngine(dev)
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = synthetic.make_config(name, device=dev)
nn = cfg["n"]
ax = torch.linspace(0, nn - 1, nn, dtype=torch.float64, device=dev)
mm = cfg["mask"].view(torch.uint8)
eng.build(cfg["points"], cfg["values"])
pore = int(cfg["mask"].sum())
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.interpolate(ax, ax, ax, mask=mm, method="rbf", k=k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print(json.dumps({"config": name, "k": k, "ms": ms, "pore_voxels_per_s": pore / ms * 1e3}))
