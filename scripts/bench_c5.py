"""Config 5 of BASELINE.json (time-resolved sweep: PTV frames at 512^3 interpolated back to back, spatial hash
rebuilt per frame) on this rank's GPU.  Frames are independent, so N GPUs run N frames at a time with no
communication (launch under torchrun; every rank takes frames rank, rank + world, ...).
usage: python scripts/bench_c5.py [frames=8] [method=sibson|idw|linear]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from ptv_interpolation_b200 import synthetic
from ptv_interpolation_b200.distributed import SlabComm
from ptv_interpolation_b200.engine import PTVEngine
from ptv_interpolation_b200.pipeline import hot_path_step

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
method = sys.argv[2] if len(sys.argv) > 2 else "sibson"
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
n = 512
mask_b = synthetic.fcc_sphere_pack_mask(n, device=dev)
mask = mask_b.view(torch.uint8)
ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)
eng = PTVEngine(dev)
comm = SlabComm(n)  # one rank owns the whole grid of its frame
out = torch.empty((3, n, n, n), dtype=torch.float32, device=dev)
ms = []
for f in range(rank, frames + world, world):  # the first one is the warm-up
    pts = synthetic.sample_pore_particles(mask_b, 5_000_000, seed=f)
    vals = synthetic.sphere_pack_flow(pts, n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = hot_path_step(eng, pts, vals, ax, ax, ax, mask, comm, method=method, k=50, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = ms[1:]
pore = int(mask_b.sum())
print(json.dumps({"config": "c5: frames of 512^3 / 5M vectors back to back, hash rebuilt per frame", "method": method,
                  "k": 50, "rank": rank, "world": world, "frames_timed": len(ms), "ms_per_frame": sum(ms) / len(ms),
                  "frames_per_s_this_gpu": 1e3 * len(ms) / sum(ms), "pore_voxels_per_s": pore * len(ms) / sum(ms) * 1e3,
                  "mean_abs_div_last": float(res.mean_abs_div)}))
