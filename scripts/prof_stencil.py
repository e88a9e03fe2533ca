"""Time the stencil kernels (fused divergence+flux, strain+vorticity) on a synthetic n^3 float32 field with a
porosity-0.4 FCC mask, for ncu / A-B runs:  python scripts/prof_stencil.py [n] [spacing] [bulk=0|1] [reps]"""
import json
import sys
import torch
sys.path.insert(0, ".")
from ptv_interpolation_b200 import synthetic
from ptv_interpolation_b200.engine import PTVEngine, set_tuning
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
h = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
bulk = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
PEAK = 6466.5
dev = torch.device("cuda", 0)
eng = PTVEngine(dev)
set_tuning(stencil_bulk=bulk)
g = torch.Generator(device=dev); g.manual_seed(0)
u, v, w = (torch.randn((n, n, n), device=dev, dtype=torch.float32, generator=g) for _ in range(3))
m = synthetic.fcc_sphere_pack_mask(n, device=dev).view(torch.uint8)
out = {"n": n, "spacing": h, "stencil_bulk": bulk}


def best(fn):
    b = 1e30
    for it in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it:
            b = min(b, e0.elapsed_time(e1))
    return b


ms = best(lambda: eng.divergence_flux(u, v, w, m, h, h, h))
out["div_flux"] = {"ms": ms, "GBps": 17.0 * n**3 / ms / 1e6, "frac": 17.0 * n**3 / ms / 1e6 / PEAK}
ms = best(lambda: eng.strain_vorticity(u, v, w, h, h, h, mask=m))
out["strain_vorticity"] = {"ms": ms, "GBps": 21.0 * n**3 / ms / 1e6, "frac": 21.0 * n**3 / ms / 1e6 / PEAK}
print(json.dumps(out))
