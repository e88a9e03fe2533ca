#!/usr/bin/env python
"""Benchmark of the PTV scattered-to-grid hot path (BASELINE.json metric: interpolated pore
voxels/sec, with HBM GB/s as a fraction of peak).

    python bench.py --gpus N --steps K --warmup W [--workload c4] [--impl reference]

A "step" is one pass of the hot path over one synthetic PTV frame that is already resident in
HBM: spatial-hash build -> fused kNN + IDW weights + solid zeroing -> masked divergence (with
z-halo exchange for N > 1) -> flux profiles + mean|div| (all-reduced for N > 1).  The grid is
sharded into z-slabs over the N ranks (strong scaling: the workload is fixed).  ``e2e`` repeats
the step through the host-buffer API: pinned host inputs are copied in and the velocity grids are
copied out inside the timed region.  ``--impl reference`` times the CPU path (the oracle port,
which calls the same SciPy cKDTree the reference calls) on a bounded sample with all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "interpolated_pore_voxels_per_sec"
UNIT = "pore voxels/s"
ALGO_BYTES_PER_VOXEL = 13.0     # 12 B (U,V,W fp32) written + 1 B mask read   
ALGO_BYTES_PER_PARTICLE = 24.0  # 6 x fp32-equivalent read once per GPU         
WORKLOADS = {
    "c1": "hex6 sphere pack 128^3, 100k vectors, IDW k=50",
    "c2": "cylinder array 256^3, 1M vectors, IDW k=50 + divergence",
    "c3": "FCC sphere pack 512^3, 5M vectors, sibson k=50",
    "c4": "dense FCC sphere pack 1024^3, 10M vectors, IDW k=50, z-slab sharded",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-voxels", type=int, default=400_000)
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# --------------------------------------------------------------------------- CPU arm
def cpu_sample_run(points, values, mask_np, n, method, k, sample_voxels, workers, canonical=False, tree=None,
                   build_s=None):
    """Time the oracle port (== the reference's own SciPy calls) on a bounded sample of the
    workload: a few whole-row z-plane windows spread through the volume, full particle set.
    Returns dict(build_s, query_s, sample_vox, sample_pore)."""
    from scipy.spatial import KDTree
    from oracle import reference_port as rp
    if tree is None:
        t0 = time.perf_counter()
        tree = KDTree(points)  # interpolator.py:132
        build_s = time.perf_counter() - t0
    rows = max(1, sample_voxels // (3 * n))
    rows = min(rows, n)
    zs = [n // 6, n // 2, (5 * n) // 6]
    ax = np.linspace(0, n - 1, n)
    query_s, nvox, npore = 0.0, 0, 0
    for z in zs:
        y0 = (n - rows) // 2
        Z, Y, X = np.meshgrid(ax[z:z + 1], ax[y0:y0 + rows], ax, indexing="ij")
        fc = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=-1)
        t0 = time.perf_counter()
        if canonical:
            dist, idx, _ = rp.knn_canonical(points, fc, k, workers=workers, tree=tree)
        else:
            dist, idx = tree.query(fc, k=k, workers=workers)  # interpolator.py:139
        if method == "sibson":
            rp.sibson_from_knn(dist, idx, values)
        else:
            rp.idw_from_knn(dist, idx, values)
        query_s += time.perf_counter() - t0
        nvox += fc.shape[0]
        npore += int(mask_np[z, y0:y0 + rows, :].sum())
    return dict(build_s=build_s, query_s=query_s, sample_vox=nvox, sample_pore=npore, tree=tree)


def cpu_throughput(r, total_vox, total_pore):
    """Whole-workload pore voxels/s implied by the sample: the reference computes every voxel
    (then zeroes solids), so time scales with ALL voxels; the tree is built once per call."""
    t_total = r["build_s"] + r["query_s"] * (total_vox / r["sample_vox"])
    return total_pore / t_total, t_total


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from ptv_interpolation_b200 import synthetic
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    cfg = synthetic.make_config(args.workload, device=dev)
    n = cfg["n"]
    pts = cfg["points"].cpu().numpy()
    vals = cfg["values"].cpu().numpy()
    mask_np = cfg["mask"].cpu().numpy()
    total_vox, total_pore = n ** 3, int(mask_np.sum())
    cores = os.cpu_count() or 1
    runs, tree, build_s = [], None, None
    for it in range(args.warmup + args.steps):
        # the tree is built (and timed) once; every step re-runs the query + weights on the sample
        r = cpu_sample_run(pts, vals, mask_np, n, cfg["method"], cfg["k"], args.cpu_sample_voxels, workers=-1,
                           tree=tree, build_s=build_s)
        tree, build_s = r["tree"], r["build_s"]
        if it >= args.warmup:
            runs.append(r)
    best = min(runs, key=lambda r: r["query_s"])
    value, t_total = cpu_throughput(best, total_vox, total_pore)
    sample = (f"{best['sample_vox']} voxels (3 z-plane windows) of {WORKLOADS[args.workload]}; cKDTree.query "
              f"workers=-1 ({cores} threads; the reference itself passes no workers= and runs 1 thread) + NumPy "
              f"weights; tree build {best['build_s']:.1f}s counted once; scaled to {total_vox} voxels")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean([r["query_s"] for r in runs])), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "grid": [n, n, n], "particles": int(len(pts)),
                   "method": cfg["method"], "k": cfg["k"], "pore_voxels": total_pore,
                   "whole_workload_s_estimate": t_total},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from ptv_interpolation_b200 import _cabi, synthetic
    from ptv_interpolation_b200.distributed import SlabComm
    from ptv_interpolation_b200.engine import PTVEngine
    from ptv_interpolation_b200.pipeline import hot_path_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the b200 arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()

    # ---- synthetic inputs, identical on every rank (same seed, same device type)
    cfg = synthetic.make_config(args.workload, device=dev)
    n, method, k = cfg["n"], cfg["method"], cfg["k"]
    comm = SlabComm(n)
    z0, z1 = comm.z0, comm.z1
    total_pore = int(cfg["mask"].sum())
    local_pore = int(cfg["mask"][z0:z1].sum())
    mask_slab = cfg["mask"][z0:z1].contiguous().view(torch.uint8)
    mask_sample_np = cfg["mask"].cpu().numpy() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    points, values = cfg["points"], cfg["values"]
    npart = points.shape[0]
    del cfg["mask"]
    ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)  # create_grid(((0,n),)*3, n)
    nzl = z1 - z0
    out = torch.empty((3, nzl, n, n), dtype=torch.float32, device=dev)
    eng = PTVEngine(dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    phase_ms = {"build": [], "interp": [], "stencils": []}

    def step(record):
        e = [ev() for _ in range(4)] if record else None
        if record:
            e[0].record()
        eng.build(points, values)
        if record:
            e[1].record()
        uvw = eng.interpolate(ax, ax, ax[z0:z1], mask=mask_slab, method=method, k=k, out=out)
        if record:
            e[2].record()
        w_below, w_above, m_above = comm.exchange_halos(uvw[2], mask_slab)
        div, stats, q_xy, q_xz, q_yz = eng.divergence_flux(uvw[0], uvw[1], uvw[2], mask_slab, 1.0, 1.0, 1.0,
                                                           w_below=w_below, w_above=w_above, mask_above=m_above)
        comm.reduce_sum_(q_xz, q_yz, stats)
        q_xy = comm.gather_planes(q_xy)
        if record:
            e[3].record()
        return e, stats

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    l0 = int(lib.ptv_launch_count())
    t_start, t_end = ev(), ev()
    t_start.record()
    evs = []
    for _ in range(args.steps):
        e, stats = step(True)
        evs.append(e)
    t_end.record()
    barrier()
    launches = int(lib.ptv_launch_count()) - l0
    clocks = sampler.stop() if sampler else None
    total_ms = t_start.elapsed_time(t_end)
    for e in evs:
        phase_ms["build"].append(e[0].elapsed_time(e[1]))
        phase_ms["interp"].append(e[1].elapsed_time(e[2]))
        phase_ms["stencils"].append(e[2].elapsed_time(e[3]))
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = total_pore / (ms_per_step * 1e-3)
    mean_abs_div = float((stats[0] / stats[1]).item())

    # ---- e2e: host buffers -> H2D -> step -> D2H, every step
    e2e = None
    if not args.no_e2e:
        hp = torch.empty(points.shape, dtype=torch.float64, pin_memory=True).copy_(points)
        hv = torch.empty(values.shape, dtype=torch.float64, pin_memory=True).copy_(values)
        hm = torch.empty(mask_slab.shape, dtype=torch.uint8, pin_memory=True).copy_(mask_slab)
        hax = torch.empty(n, dtype=torch.float64, pin_memory=True).copy_(ax)
        hout = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)
        hstats = torch.empty(2, dtype=torch.float64, pin_memory=True)
        dp, dv, dm, dax = (torch.empty_like(t, device=dev) for t in (hp, hv, hm, hax))

        def e2e_step():
            # the particle table enters the box once (rank 0, pinned host -> HBM) and is replicated over
            # NVLink; every rank copies in its own mask slab and copies out its own U,V,W slab
            if rank == 0 or world == 1:
                dp.copy_(hp, non_blocking=True)
                dv.copy_(hv, non_blocking=True)
            if world > 1:
                dist.broadcast(dp, 0)
                dist.broadcast(dv, 0)
            dm.copy_(hm, non_blocking=True)
            dax.copy_(hax, non_blocking=True)
            eng.build(dp, dv)
            # z-chunked search with the device->host copy of finished chunks overlapped on a second stream
            uvw = eng.interpolate_to_host(dax, dax, dax[z0:z1], hout, mask=dm, dev_out=out, method=method, k=k)
            w_below, w_above, m_above = comm.exchange_halos(uvw[2], dm)
            div, st, _, _, _ = eng.divergence_flux(uvw[0], uvw[1], uvw[2], dm, 1.0, 1.0, 1.0, w_below=w_below,
                                                   w_above=w_above, mask_above=m_above)
            comm.reduce_sum_(st)
            hstats.copy_(st, non_blocking=True)
            torch.cuda.current_stream().wait_stream(eng.copy_stream)  # the step ends when U,V,W are on the host

        e2e_step()
        barrier()
        n_e2e = max(1, min(args.steps, 3))
        a, b = ev(), ev()
        a.record()
        for _ in range(n_e2e):
            e2e_step()
        b.record()
        barrier()
        t2 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item()) / n_e2e
        h2d = hp.numel() * 8 + hv.numel() * 8 + hm.numel() + hax.numel() * 8  # rank 0; other ranks: mask slab + axes
        d2h = hout.numel() * 4 + 16
        e2e = {"value": total_pore / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "note": "rank-0 bytes; pinned host buffers; particles enter through rank 0 and are broadcast over "
                       "NVLink when N > 1; result = U,V,W slab + (sum|div|, n_fluid)"}
        del hp, hv, hm, hout, dp, dv, dm

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (fused kNN + weights), timed live with CUDA events
    peak, peak_src = measured_peak()
    knn_ms = float(np.mean(phase_ms["interp"]))
    algo_bytes = ALGO_BYTES_PER_VOXEL * nzl * n * n + ALGO_BYTES_PER_PARTICLE * npart
    achieved = algo_bytes / (knn_ms * 1e-3) / 1e9
    used_stream = eng.knn_stats()["used_stream"]
    roofline = {"kernel": "knn_stream_kernel (+ knn_interp_kernel on handed-over tiles)" if used_stream
                else "knn_interp_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu capture of
                # this workload (profiles/r01_v4_knn_stream_c4_ncu_metrics.csv); null for other workloads
                "traffic": 21.52e9 if (args.workload == "c4" and world == 1 and used_stream) else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": knn_ms,
                "note": "kNN selection is SM-issue bound, not HBM bound (DESIGN.md); the HBM-bound kernels are "
                        "listed under roofline_other"}
    # What actually bounds that kernel: warp-instruction issue.  Instructions per launch come from the same
    # committed ncu capture (smsp__inst_executed.sum); the time is this run's; peak = 148 SMs x 4 schedulers x
    # one warp instruction per clock at the maximum SM clock.
    roofline_issue = None
    if args.workload == "c4" and world == 1 and used_stream:
        issue_peak = 148 * 4 * 1.965e9
        issue_ach = 2.288e11 / (knn_ms * 1e-3)
        roofline_issue = {"kernel": "knn_stream_kernel", "bound": "sm_issue", "achieved": issue_ach / 1e9,
                          "peak": issue_peak / 1e9, "unit": "G warp-inst/s", "frac": issue_ach / issue_peak,
                          "warp_inst_per_launch": 2.288e11,
                          "source": "profiles/r01_v4_knn_stream_c4_ncu_metrics.csv (smsp__inst_executed.sum; "
                                    "smsp__issue_active 56.8 % in that capture)"}
    st_ms = float(np.mean(phase_ms["stencils"]))
    st_bytes = 17.0 * nzl * n * n  # fused divergence + flux + statistics: 12 B u,v,w + 1 B mask read, 4 B div written
    roofline_other = [{"kernel": "div_flux_kernel (1 launch + halo/reduce)", "bound": "hbm",
                       "achieved": st_bytes / (st_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                       "frac": st_bytes / (st_ms * 1e-3) / 1e9 / peak, "ms": st_ms},
                      {"kernel": "hash build (10 launches)", "bound": "hbm",
                       "achieved": 112.0 * npart / (float(np.mean(phase_ms["build"])) * 1e-3) / 1e9, "peak": peak,
                       "unit": "GB/s", "frac": 112.0 * npart / (float(np.mean(phase_ms["build"])) * 1e-3) / 1e9 / peak,
                       "ms": float(np.mean(phase_ms["build"]))}]

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_sample_run(points.cpu().numpy(), values.cpu().numpy(), mask_sample_np, n, method, k,
                           args.cpu_sample_voxels, workers=1)
        v, t_total = cpu_throughput(r, n ** 3, total_pore)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                        "host_cores_available": os.cpu_count(),
                        "sample": f"{r['sample_vox']} voxels (3 z-plane windows) of the same workload through the "
                                  f"oracle port as the reference runs it (cKDTree.query workers=1 + NumPy weights): "
                                  f"{r['query_s']:.1f}s query + {r['build_s']:.1f}s tree build; scaled to {n**3} "
                                  f"voxels -> {t_total:.0f}s per frame"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 distances/weights, f32 output", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "grid": [n, n, n], "particles": int(npart),
                   "method": method, "k": k, "pore_voxels": total_pore, "porosity": total_pore / n ** 3,
                   "parallelism": f"z-slab x{world}", "slab_planes_rank0": nzl,
                   "l2_policy": "inputs_larger_than_L2" if nzl * n * n * 13 > 126e6 else "small_workload_fits_L2",
                   "mask_skip": True, "all_voxels_per_sec": n ** 3 / (ms_per_step * 1e-3),
                   "phase_ms_rank0": {p: float(np.mean(v)) for p, v in phase_ms.items()},
                   "mean_abs_div": mean_abs_div},
        "roofline": roofline, "roofline_issue": roofline_issue, "roofline_other": roofline_other,
        "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
