"""Timings of method='linear' (Delaunay tetrahedron + barycentric weights, csrc/delaunay_linear.cu) on one
GPU -> JSON line.  Usage: python scripts/bench_linear.py [c1 c2 c3 c4]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from ptv_interpolation_b200 import synthetic
from ptv_interpolation_b200.engine import PTVEngine, set_tuning

dev = torch.device("cuda", 0)
eng = PTVEngine(dev)
out = {}
set_tuning(stats=1)
names = [a for a in sys.argv[1:] if "=" not in a] or ["c1", "c2"]
for a in sys.argv[1:]:
    if "=" in a:
        set_tuning(**{a.split("=")[0]: float(a.split("=")[1])})
for name in names:
    cfg = synthetic.make_config(name, device=dev)
    n = cfg["n"]
    ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)
    mask = cfg["mask"].view(torch.uint8)
    eng.build(cfg["points"], cfg["values"])
    res = torch.empty((3, n, n, n), dtype=torch.float32, device=dev)
    best = 1e30
    first = 0.0
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.interpolate(ax, ax, ax, mask=mask, method="linear", out=res)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
        else:
            first = e0.elapsed_time(e1)  # includes building the hull-candidate list for this hash
    pore = int(cfg["mask"].sum())
    out[f"linear_{name}"] = {"grid": n, "particles": int(cfg["points"].shape[0]), "pore_voxels": pore, "ms": best, "first_call_ms": first,
                             "pore_voxels_per_s": pore / best * 1e3, "stats": eng.linear_stats()}
    if n <= 256:  # the same through the drop-in API: host DataFrame in, host arrays out (wall clock)
        import time
        import numpy as np
        import pandas as pd
        from ptv_interpolation_b200 import interpolator as gi
        p_h, v_h, m_h = cfg["points"].cpu().numpy(), cfg["values"].cpu().numpy(), cfg["mask"].cpu().numpy()
        df = pd.DataFrame({"x": p_h[:, 0], "y": p_h[:, 1], "z": p_h[:, 2], "u": v_h[:, 0], "v": v_h[:, 1], "w": v_h[:, 2]})
        grid, _ = gi.create_grid(((0, n), (0, n), (0, n)), n)
        gi.interpolate_field(df, grid, method="linear", mask=m_h)
        t0 = time.perf_counter()
        gi.interpolate_field(df, grid, method="linear", mask=m_h)
        out[f"linear_{name}"]["e2e_interpolate_field_ms"] = (time.perf_counter() - t0) * 1e3
    print(name, out[f"linear_{name}"], file=sys.stderr, flush=True)
    del cfg, mask, res
print(json.dumps(out))
