#!/usr/bin/env python
"""Host <-> device transfer probes behind the end-to-end numbers (one process per GPU under torchrun):

* D2H sink: every rank drains its z-slab of a 1024^3 x 3 float32 result (12.9 GB / N per rank) from HBM into
  pinned host memory, all ranks at once -> aggregate GB/s the host can sink (the floor of the e2e frame time).
* rank 0 only: pageable -> device paths for the inputs (pandas column extraction, torch pageable copy,
  staged pinned chunks, cudaHostRegister in place), pinned H2D.

    python scripts/host_io_probe.py [--out profiles/r02_host_sink.json]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/host_io_probe.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--bind", action="store_true", help="pin every rank to the CPUs next to its GPU first")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = {}
    if a.bind:
        from ptv_interpolation_b200 import hostmem
        numa = hostmem.bind_host_to_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.n
    nzl = n // world
    res = {"world": world, "grid": n, "numa_binding_rank0": numa}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- D2H sink
    d = torch.zeros((3, nzl, n, n), dtype=torch.float32, device=dev)
    h = torch.empty((3, nzl, n, n), dtype=torch.float32, pin_memory=True)
    h.copy_(d)
    times = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    nbytes = 3 * n ** 3 * 4
    res["d2h_sink"] = {"bytes_total": nbytes, "best_s": min(times), "aggregate_gb_s": nbytes / min(times) / 1e9,
                       "per_rank_gb_s": nbytes / world / min(times) / 1e9}
    # ---- H2D, all ranks at once (mask slab sized)
    hm = torch.empty((nzl, n, n), dtype=torch.uint8, pin_memory=True)
    dm = torch.empty((nzl, n, n), dtype=torch.uint8, device=dev)
    dm.copy_(hm)
    barrier()
    t0 = time.perf_counter()
    dm.copy_(hm, non_blocking=True)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["h2d_pinned_mask"] = {"bytes_total": n ** 3, "aggregate_gb_s": n ** 3 / float(t.item()) / 1e9}
    del d, h, hm, dm

    if rank == 0:
        import pandas as pd
        from ptv_interpolation_b200 import hostmem
        npart = 10_000_000
        rng = np.random.default_rng(0)
        cols = {c: rng.random(npart) for c in "xyzuvw"}
        df = pd.DataFrame(cols)
        t0 = time.perf_counter()
        pts = df[["x", "y", "z"]].values
        t_vals = time.perf_counter() - t0
        res["pandas_values_240MB_s"] = t_vals
        pts = np.ascontiguousarray(pts)
        mask = np.zeros((n // 2, n, n), dtype=np.bool_)  # 0.5 GB of pageable memory
        mask[::3] = True
        probes = {}
        for name, arr in (("points_240MB", pts), ("mask_512MB", mask)):
            nb = arr.nbytes
            t0 = time.perf_counter()
            x = torch.from_numpy(arr.view(np.uint8) if arr.dtype == np.bool_ else arr).to(dev)
            torch.cuda.synchronize()
            t_pageable = time.perf_counter() - t0
            hostmem.stage_to_device(arr, dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            y = hostmem.stage_to_device(arr, dev)
            torch.cuda.synchronize()
            t_staged = time.perf_counter() - t0
            ok = bool(torch.equal(x.view(torch.uint8).reshape(-1), y.view(torch.uint8).reshape(-1)))
            # pin the caller's pages in place, DMA, unpin
            rt = torch.cuda.cudart()
            flat = arr.reshape(-1).view(np.uint8)
            t0 = time.perf_counter()
            rc = rt.cudaHostRegister(flat.ctypes.data, nb, 0)
            t_reg = time.perf_counter() - t0
            t_dma = t_unreg = None
            if int(rc) == 0:
                z = torch.empty(nb, dtype=torch.uint8, device=dev)
                t0 = time.perf_counter()
                z.copy_(torch.from_numpy(flat), non_blocking=True)
                torch.cuda.synchronize()
                t_dma = time.perf_counter() - t0
                t0 = time.perf_counter()
                rt.cudaHostUnregister(flat.ctypes.data)
                t_unreg = time.perf_counter() - t0
            probes[name] = {"bytes": nb, "pageable_to_s": t_pageable, "staged_chunks_s": t_staged, "staged_ok": ok,
                            "host_register_s": t_reg, "registered_dma_s": t_dma, "host_unregister_s": t_unreg,
                            "pageable_gb_s": nb / t_pageable / 1e9, "staged_gb_s": nb / t_staged / 1e9}
        res["pageable_inputs_rank0"] = probes
        res["torch_cpu_threads"] = torch.get_num_threads()
        res["host_cores"] = os.cpu_count()
        print(json.dumps(res))
        if a.out:
            with open(os.path.join(ROOT, a.out), "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
