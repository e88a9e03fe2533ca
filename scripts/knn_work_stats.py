#!/usr/bin/env python
"""Work counters of the streaming kNN kernel on one benchmark config (tuning "stats" = 1):
voxel-candidate pairs per pore voxel in the histogram and classification passes, exact keys,
list entries, fail-list rate, plus the kernel time without the counters.

    python scripts/knn_work_stats.py [c4] [--planes 128] [--out profiles/r02_knn_work_c4.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ptv_interpolation_b200 import synthetic  # noqa: E402
from ptv_interpolation_b200.engine import PTVEngine, set_tuning  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="c4")
    ap.add_argument("--planes", type=int, default=0, help="only the first N z-planes of the grid (0 = all)")
    ap.add_argument("--out", default="")
    ap.add_argument("--tuning", default="", help="comma-separated key=value tuning overrides")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = synthetic.make_config(a.workload, device=dev)
    n, method, k = cfg["n"], cfg["method"], cfg["k"]
    nz = a.planes or n
    z0 = (n - nz) // 2
    mask = cfg["mask"][z0:z0 + nz].contiguous().view(torch.uint8)
    ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)
    eng = PTVEngine(dev)
    for kv in filter(None, a.tuning.split(",")):
        key, val = kv.split("=")
        set_tuning(**{key: float(val)})
    eng.build(cfg["points"], cfg["values"])
    out = torch.empty((3, nz, n, n), dtype=torch.float32, device=dev)
    pore = int(mask.sum())

    def run():
        eng.interpolate(ax, ax, ax[z0:z0 + nz], mask=mask, method=method, k=k, out=out)

    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    set_tuning(stats=1)
    run()
    torch.cuda.synchronize()
    st = eng.knn_stats()
    set_tuning(stats=0)
    w = st["work"]
    vox = max(w["voxels"], 1)
    res = {"workload": a.workload, "planes": nz, "pore_voxels": pore, "method": method, "k": k, "kernel_ms": ms,
           "pore_voxels_per_s": pore / (ms * 1e-3), "stats": st,
           "per_voxel": {key: w[key] / vox for key in ("pairs_histogram", "pairs_classify", "exact_keys", "list_entries")},
           "voxels_per_round": vox / max(w["rounds"], 1), "checksum": float(out.double().sum().item())}
    print(json.dumps(res))
    if a.out:
        with open(os.path.join(ROOT, a.out), "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
