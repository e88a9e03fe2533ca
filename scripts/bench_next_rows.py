"""Timings of the SURVEY 8f 'next' rows on one GPU -> JSON line (profiles/rNN_next_rows.json)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from ptv_interpolation_b200 import synthetic
from ptv_interpolation_b200.engine import PTVEngine

PEAK = 6466.5
dev = torch.device("cuda", 0)
eng = PTVEngine(dev)
out = {}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r


# a7 / a8: sibson (c3: 512^3, 5M vectors, k=50) and local RBF (c2: 256^3, 1M vectors, k=20) on pore voxels
for name, method, k in (("c3", "sibson", 50), ("c2", "rbf", 20), ("c2", "nearest", 1)):
    cfg = synthetic.make_config(name, device=dev)
    nn = cfg["n"]
    axx = torch.linspace(0, nn - 1, nn, dtype=torch.float64, device=dev)
    mm = cfg["mask"].view(torch.uint8)
    eng.build(cfg["points"], cfg["values"])
    ms, _ = timed(lambda: eng.interpolate(axx, axx, axx, mask=mm, method=method, k=k), reps=2)
    pore = int(cfg["mask"].sum())
    out[f"{method}_{name}"] = {"k": k, "grid": nn, "ms": ms, "pore_voxels_per_s": pore / ms * 1e3}
    del cfg, mm

# N1: outlier filter, 10M particles of the c4 cloud, k = 25
cfg = synthetic.make_config("c4", device=dev)
eng.build(cfg["points"], cfg["values"])
ms, (keep, kth) = timed(lambda: eng.outlier_filter(k=25, threshold=3.0), reps=2)
out["N1_outlier_filter"] = {"particles": int(cfg["points"].shape[0]), "k": 25, "ms": ms,
                            "particles_per_s": cfg["points"].shape[0] / ms * 1e3, "kept_fraction": float(keep.float().mean())}
n = cfg["n"]
mask = cfg["mask"].view(torch.uint8)
del cfg
# N3: strain rate + vorticity, 1024^3 float32 (13 B read + 8 B written per voxel)
g = torch.Generator(device=dev); g.manual_seed(1)
u, v, w = (torch.randn((n, n, n), device=dev, dtype=torch.float32, generator=g) for _ in range(3))
ms, _ = timed(lambda: eng.strain_vorticity(u, v, w, 1.0, 1.0, 1.0, mask=mask))
out["N3_strain_vorticity_1024"] = {"ms": ms, "GBps": 21.0 * n**3 / ms / 1e6, "frac_of_measured_peak": 21.0 * n**3 / ms / 1e6 / PEAK}
ms, _ = timed(lambda: eng.strain_vorticity(u, v, w, 1.25, 0.75, 2.0, mask=mask))
out["N3_strain_vorticity_1024_nonunit"] = {"ms": ms, "GBps": 21.0 * n**3 / ms / 1e6}
del u, v, w
# N2: LSQR on the 512^3 FCC pack (divergence of a random field): time per iteration and a capped solve
n2 = 512
m2 = synthetic.fcc_sphere_pack_mask(n2, device=dev).view(torch.uint8)
div = torch.randn((n2, n2, n2), device=dev, dtype=torch.float32, generator=g) * m2
t0 = time.perf_counter()
phi, info = eng.poisson_lsqr(div, m2, 1.0, 1.0, 1.0, iter_lim=200)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out["N2_lsqr_512"] = {"fluid_voxels": int(m2.sum()), "iterations": info["itn"], "istop": info["istop"],
                      "ms_per_iteration": dt * 1e3 / max(info["itn"], 1),
                      "GBps": (2 * 25.0 + 40.0) * n2**3 * info["itn"] / dt / 1e9}
ms, _ = timed(lambda: eng.projection_correct(div, div, div, phi, m2, 1.0, 1.0, 1.0))
out["N2_correction_512"] = {"ms": ms, "GBps": (12 + 8 + 1 + 12) * n2**3 / ms / 1e6}
# a2 / a3: mask resampling gather (2 B/voxel) and boundary-voxel extraction at 1024^3
del div, phi, m2
n = 1024
mraw = synthetic.fcc_sphere_pack_mask(n, device=dev).view(torch.uint8)
ident = torch.arange(n, device=dev, dtype=torch.int32)
ms, _ = timed(lambda: eng.mask_gather(mraw, ident, ident, ident))
out["a2_mask_gather_1024"] = {"ms": ms, "GBps": 2.0 * n**3 / ms / 1e6, "frac_of_measured_peak": 2.0 * n**3 / ms / 1e6 / PEAK}
half = torch.arange(0, n, 2, device=dev, dtype=torch.int32)
ms, _ = timed(lambda: eng.mask_gather(mraw, half, half, half))
out["a2_mask_gather_1024_to_512"] = {"ms": ms}
for th in (1, 2):
    walls = []
    for rep in range(3):  # first call allocates the workspace; the later ones are what a frame loop pays
        t0 = time.perf_counter()
        idx = eng.boundary_voxels(mraw, thickness=th)
        torch.cuda.synchronize()
        walls.append((time.perf_counter() - t0) * 1e3)
        cnt = int(idx.numel())
        del idx
    out[f"a3_boundary_voxels_1024_t{th}"] = {"ms_wall_first": walls[0], "ms_wall": min(walls[1:]), "count": cnt}
print(json.dumps(out))
