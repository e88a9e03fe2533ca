#!/usr/bin/env python
"""Benchmark of the PTV scattered-to-grid hot path (BASELINE.json metric: interpolated pore
voxels/sec, with HBM GB/s as a fraction of peak).

    python bench.py --gpus N --steps K --warmup W [--workload c4|c5|c1|c2|c3] [--impl reference]

A "step" is one pass of the hot path over one synthetic PTV frame that is already resident in
HBM: spatial-hash build -> fused kNN + IDW weights + solid zeroing -> masked divergence (with
z-halo exchange for N > 1) -> flux profiles + mean|div| (all-reduced for N > 1).  The grid is
sharded into z-slabs over the N ranks (strong scaling: the workload is fixed).

``e2e`` is the same frame through the reference-facing plugin call: every rank calls
``interpolator.interpolate_field(df, grid, method=..., mask=...)`` with a pandas DataFrame and NumPy
arrays in pageable host memory and gets NumPy U, V, W back -- host->device and device->host copies
inside the timed region (wall clock around the call, max over ranks).  ``e2e_engine`` keeps the
engine-level number (pinned buffers allocated once, particles broadcast from rank 0).

Outside the timed regions the line also carries: ``parity_check`` (sampled voxels of this rank's slab and
the slab-boundary divergence plane against the oracle), ``cpu_baseline`` (the oracle port as the reference
runs it, 1 thread, bounded sample), ``cpu_rows`` (the reference's multiprocessing RBF mode and the
workers=-1 best case, next to the GPU numbers for the same sample) and the all-voxel (no ``mask=``) time.

``--workload c5`` is the time-resolved sweep (64 frames at 512^3, hash rebuilt per frame) in both
frame-parallel and slab-parallel form.  ``--impl reference`` times the CPU path (the oracle port, which
calls the same SciPy cKDTree the reference calls) on a bounded sample with all host cores.
"""
from __future__ import annotations

import argparse
import hashlib
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
    "c3": "FCC sphere pack 512^3, 5M vectors + wall particles, sibson k=50",
    "c4": "dense FCC sphere pack 1024^3, 10M vectors, IDW k=50, z-slab sharded",
    "c5": "time-resolved sweep: 64 PTV frames at 512^3 (config-3 frames, sibson k=50), hash rebuilt per frame",
}
KNN_KERNEL_SOURCE = os.path.join(ROOT, "ptv_interpolation_b200", "csrc", "knn_duo.cu")
KNN_PROFILE = os.path.join(ROOT, "profiles", "r02_knn_duo_c4_ncu.json")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-rows", action="store_true", help="skip the RBF process-pool / workers=-1 CPU rows")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-all-voxel", action="store_true")
    ap.add_argument("--halo-overlap", action="store_true",
                    help="interpolate the slab's boundary planes first and exchange halos behind the interior (3 launches)")
    ap.add_argument("--no-halo-overlap", action="store_true", help=argparse.SUPPRESS)  # the default since round 2
    ap.add_argument("--cpu-sample-voxels", type=int, default=400_000)
    ap.add_argument("--parity-voxels", type=int, default=4000)
    ap.add_argument("--c5-frames", type=int, default=64)
    ap.add_argument("--rbf-rows-from", default="", help=argparse.SUPPRESS)  # internal: CPU RBF pool row in a clean process
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def knn_profile(workload, world):
    """Per-launch counters of the dominant kernel from the committed ncu capture -- used only if the capture
    was taken from the kernel source that is compiled now (content hash), for this workload on one GPU."""
    if workload != "c4" or world != 1 or not os.path.exists(KNN_PROFILE):
        return None
    try:
        prof = json.load(open(KNN_PROFILE))
        with open(KNN_KERNEL_SOURCE, "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != prof.get("kernel_source_sha256"):
                return None
        return prof
    except Exception:
        return None


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


def core_config(workload, n, npart, method, k, pore):
    """The keys both arms report identically (the driver compares them)."""
    return {"workload": WORKLOADS[workload], "grid": [n, n, n], "particles": int(npart), "method": method, "k": int(k),
            "pore_voxels": int(pore)}


# --------------------------------------------------------------------------- CPU arm
def sample_windows(n, sample_voxels):
    """A few whole-row z-plane windows spread through the volume: [(z, y0, rows)]."""
    rows = min(max(1, sample_voxels // (3 * n)), n)
    return [(z, (n - rows) // 2, rows) for z in (n // 6, n // 2, (5 * n) // 6)]


def window_coords(n, z, y0, rows):
    ax = np.linspace(0, n - 1, n)
    Z, Y, X = np.meshgrid(ax[z:z + 1], ax[y0:y0 + rows], ax, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=-1)


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
    query_s, nvox, npore = 0.0, 0, 0
    for z, y0, rows in sample_windows(n, sample_voxels):
        fc = window_coords(n, z, y0, rows)
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


def rbf_pool_row(npz_path):
    """CPU row 2: the reference's own multiprocessing mode -- method='rbf',
    n_jobs=<all host cores> (interpolator.py:173-182, test_parallel.py:24) -- through the oracle port on a
    bounded sample of the workload.  Runs in a process of its own (no CUDA context to fork)."""
    from oracle import reference_port as rp
    d = np.load(npz_path)
    pts, vals, n, total_vox, total_pore = d["points"], d["values"], int(d["n"]), int(d["total_vox"]), int(d["total_pore"])
    cores = os.cpu_count() or 1
    res = {}
    for tag, wins in (("a", d["windows"][:1]), ("b", d["windows"])):
        fc = np.concatenate([window_coords(n, int(z), int(y0), int(rows)) for z, y0, rows in wins], 0)
        grid = (fc[:, 0].reshape(1, 1, -1), fc[:, 1].reshape(1, 1, -1), fc[:, 2].reshape(1, 1, -1))
        t0 = time.perf_counter()
        rp.interpolate_field(pts, vals, grid, method="rbf", rbf_neighbors=20, n_jobs=cores)
        res[tag] = (time.perf_counter() - t0, fc.shape[0])
    # two sample sizes separate the per-call cost (tree build, pickling the interpolator to every worker)
    # from the per-voxel cost
    (ta, na), (tb, nb) = res["a"], res["b"]
    per_vox = max((tb - ta) / max(nb - na, 1), 1e-12)
    fixed = max(ta - per_vox * na, 0.0)
    t_total = fixed + per_vox * total_vox
    print(json.dumps({"value": total_pore / t_total, "unit": UNIT, "cores": cores, "kind": "port",
                      "mode": "method='rbf', rbf_neighbors=20, n_jobs=%d (ProcessPoolExecutor)" % cores,
                      "sample": f"{na} and {nb} voxels of the same workload: {ta:.1f}s and {tb:.1f}s -> "
                                f"{fixed:.1f}s per call + {per_vox * 1e6:.1f} us/voxel; scaled to {total_vox} voxels "
                                f"-> {t_total:.0f}s per frame",
                      "voxels_per_sec_all": 1.0 / per_vox}))
    return 0


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from ptv_interpolation_b200 import synthetic
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    wl = "c3" if args.workload == "c5" else args.workload
    cfg = synthetic.make_config(wl, device=dev)
    n = cfg["n"]
    pts = cfg["points"].cpu().numpy()
    vals = cfg["values"].cpu().numpy()
    mask_np = cfg["mask"].cpu().numpy()
    total_vox, total_pore = n ** 3, int(mask_np.sum())
    frames = args.c5_frames if args.workload == "c5" else 1
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
    value, t_total = cpu_throughput(best, total_vox, total_pore)  # per frame; frames are independent
    sample = (f"{best['sample_vox']} voxels (3 z-plane windows) of {WORKLOADS[args.workload]}; cKDTree.query "
              f"workers=-1 ({cores} threads; the reference itself passes no workers= and runs 1 thread) + NumPy "
              f"weights; tree build {best['build_s']:.1f}s counted once per frame; scaled to {total_vox} voxels")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean([r["query_s"] for r in runs])), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": core_config(args.workload, n, len(pts), cfg["method"], cfg["k"], total_pore * frames),
        "details": {"whole_workload_s_estimate": t_total * frames},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- GPU arm
def _dist_env():
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def parity_check(eng, points_np, values_np, tree, ax_np, z0, z1, mask_slab_np, out, div, halos, nvox, seed):
    """Outside the timed region: `nvox` random pore voxels of this rank's slab against the oracle
    (canonical cKDTree neighbours + the reference's weights, interpolator.py:126-155 / 95-123), and the
    divergence of the slab's last plane -- the one that needs the upper neighbour's halo -- against
    physics.py:6-53 evaluated on the GPU's own field."""
    import torch
    from oracle import reference_port as rp
    method, k = eng["method"], eng["k"]
    rng = np.random.default_rng(seed)
    pore = np.flatnonzero(mask_slab_np.reshape(-1))
    res = {"voxels": 0, "max_rel_err": 0.0, "tolerance": 1e-5}
    if len(pore):
        sel = rng.choice(pore, size=min(nvox, len(pore)), replace=False)
        ny, nx = mask_slab_np.shape[1:]
        zz, rem = np.divmod(sel, ny * nx)
        yy, xx = np.divmod(rem, nx)
        q = np.stack([ax_np[xx], ax_np[yy], ax_np[z0 + zz]], -1)
        d, i, _ = rp.knn_canonical(points_np, q, k, workers=-1, tree=tree)
        ref = (rp.sibson_from_knn(d, i, values_np) if method == "sibson" else rp.idw_from_knn(d, i, values_np)).T
        got = out.reshape(3, -1)[:, torch.from_numpy(sel).to(out.device)].double().cpu().numpy()
        scale = np.sqrt((values_np ** 2).mean(0))[:, None]
        err = np.abs(got - ref) / np.maximum(np.abs(ref), scale)
        res.update(voxels=int(len(sel)), max_rel_err=float(err.max()))
    # slab-boundary divergence plane (z1-1): own planes z1-2, z1-1 plus the halo plane z1 (w and mask only)
    nzl = z1 - z0
    if div is not None and nzl >= 2:
        w_below, w_above, m_above = halos
        top = 2 if w_above is not None else 1
        u = np.zeros((top + 1,) + mask_slab_np.shape[1:])
        v, w, m = np.zeros_like(u), np.zeros_like(u), np.zeros(u.shape, dtype=bool)
        o = out[:, nzl - 2:nzl].double().cpu().numpy()
        u[:2], v[:2], w[:2] = o[0], o[1], o[2]
        m[:2] = mask_slab_np[nzl - 2:nzl] != 0
        if w_above is not None:
            w[2] = w_above.double().cpu().numpy()
            m[2] = m_above.cpu().numpy() != 0
        # the mini-volume starts at slab plane nzl-2: its first plane has no lower neighbour in the mini-volume,
        # so only the middle plane (slab plane nzl-1) is compared when a halo plane exists
        dref = rp.compute_consistent_divergence(u, v, w, m, 1.0, 1.0, 1.0)
        if w_above is not None:
            dgot = div[nzl - 1].double().cpu().numpy()
            res["divergence_plane"] = {"plane": int(z1 - 1), "uses_halo": True,
                                       "max_abs_diff": float(np.abs(dgot - dref[1].astype(np.float32)).max())}
        else:
            dgot = div[nzl - 1].double().cpu().numpy()
            res["divergence_plane"] = {"plane": int(z1 - 1), "uses_halo": False,
                                       "max_abs_diff": float(np.abs(dgot - dref[1].astype(np.float32)).max())}
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist
    from ptv_interpolation_b200 import _cabi, synthetic
    from ptv_interpolation_b200 import interpolator as gi
    from ptv_interpolation_b200.distributed import SlabComm
    from ptv_interpolation_b200.engine import PTVEngine, set_tuning
    from ptv_interpolation_b200.pipeline import hot_path_step

    world, rank, local = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the b200 arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from ptv_interpolation_b200 import hostmem
    numa = hostmem.bind_host_to_device(local)  # node-local pinned buffers for the host<->device copies
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()
    if args.workload == "c5":
        return run_c5(args, dev, world, rank, local, lib)

    # ---- synthetic inputs, identical on every rank (same seed, same device type)
    cfg = synthetic.make_config(args.workload, device=dev)
    n, method, k = cfg["n"], cfg["method"], cfg["k"]
    # slabs balanced by pore voxels per plane (solid voxels cost nothing); every rank computes the same cuts
    plane_pore = cfg["mask"].sum(dim=(1, 2)).cpu().tolist()
    comm = SlabComm(n, plane_weights=plane_pore)
    z0, z1 = comm.z0, comm.z1
    total_pore = int(cfg["mask"].sum())
    mask_slab = cfg["mask"][z0:z1].contiguous().view(torch.uint8)
    mask_slab_np = mask_slab.cpu().numpy()
    mask_sample_np = cfg["mask"].cpu().numpy() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    points, values = cfg["points"], cfg["values"]
    points_np, values_np = points.cpu().numpy(), values.cpu().numpy()
    npart = points.shape[0]
    del cfg["mask"]
    ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)  # create_grid(((0,n),)*3, n)
    ax_np = np.linspace(0, n - 1, n)
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
    keep = {}

    def step(record):
        e = [ev() for _ in range(4)] if record else None
        marks = {"built": 1, "interpolated": 2}
        if record:
            e[0].record()
        res = hot_path_step(eng, points, values, ax, ax, ax, mask_slab, comm, method=method, k=k, out=out,
                            mark=(lambda label: e[marks[label]].record()) if record else None,
                            overlap_halos=args.halo_overlap and not args.no_halo_overlap)
        if record:
            e[3].record()
        keep.update(div=res.div, res=res)
        return e, (res.mean_abs_div * res.n_fluid, res.n_fluid)

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

    # ---- work counters of the streaming kernel (one extra launch with tuning "stats" = 1)
    set_tuning(stats=1)
    eng.interpolate(ax, ax, ax[z0:z1], mask=mask_slab, method=method, k=k, out=out)
    torch.cuda.synchronize()
    kst = eng.knn_stats()
    set_tuning(stats=0)
    wk = kst["work"]
    work = None
    if kst["used_stream"] and wk["voxels"] > 0:
        work = {"candidates_per_voxel_histogram_pass": wk["pairs_histogram"] / wk["voxels"],
                "candidates_per_voxel_classify_pass": wk["pairs_classify"] / wk["voxels"],
                "exact_keys_per_voxel": wk["exact_keys"] / wk["voxels"],
                "list_entries_per_voxel": wk["list_entries"] / wk["voxels"],
                "voxels_per_warp_pass": wk["voxels"] / max(wk["rounds"], 1),
                "heap_tiles_redone": kst["tiles_failed"], "fail_reasons": kst["fail_reasons"]}

    # ---- parity of what was just timed, against the oracle (outside the timed region)
    parity = None
    tree = None
    if not args.no_parity:
        from scipy.spatial import KDTree
        tree = KDTree(points_np)
        halos = comm.exchange_halos(out[2], mask_slab)  # the planes the timed step used (again, for the check)
        parity = parity_check(dict(method=method, k=k), points_np, values_np, tree, ax_np, z0, z1, mask_slab_np, out,
                              keep.get("div"), halos, args.parity_voxels, seed=100 + rank)
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, parity)
            parity = {"voxels": int(sum(g["voxels"] for g in gathered)),
                      "max_rel_err": float(max(g["max_rel_err"] for g in gathered)), "tolerance": 1e-5,
                      "divergence_planes": [g.get("divergence_plane") for g in gathered], "ranks": world}
        parity["ok"] = bool(parity["max_rel_err"] <= parity["tolerance"])

    # ---- e2e: the plugin call, NumPy / pandas in pageable host memory in, NumPy out, every step
    e2e = e2e_engine = None
    all_voxel_ms = None
    if not args.no_e2e:
        import pandas as pd
        df = pd.DataFrame({"x": points_np[:, 0], "y": points_np[:, 1], "z": points_np[:, 2],
                           "u": values_np[:, 0], "v": values_np[:, 1], "w": values_np[:, 2]})
        (X, Y, Z), _axes = gi.create_grid(((0, n), (0, n), (0, n)), n)
        slab_grid = (X[z0:z1], Y[z0:z1], Z[z0:z1])  # this rank's z-slab of the create_grid() mesh
        mask_bool = mask_slab_np.view(np.bool_)
        kw = dict(method=method, idw_neighbors=k, sibson_neighbors=k)
        del out
        torch.cuda.empty_cache()

        def plugin_call(mask):
            U, V, W = gi.interpolate_field(df, slab_grid, mask=mask, device=dev, shard_inputs=world > 1, **kw)
            return float(W[0, 0, 0])  # the result is host memory the caller can read

        plugin_call(mask_bool)  # first call allocates the pinned pools
        plugin_call(mask_bool)
        barrier()
        n_e2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            plugin_call(mask_bool)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        t2 = torch.tensor([(t1 - t0) * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item()) / n_e2e
        h2d = -(-npart // world) * 48 + mask_slab_np.size + 3 * n * 8
        d2h = 3 * nzl * n * n * 4
        e2e = {"value": total_pore / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
               "via": "interpolator.interpolate_field(df, grid, method=%r, mask=mask)" % method,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "timer": "wall clock, max over ranks",
               "note": "per-rank bytes; pandas DataFrame + NumPy mask in pageable host memory -> NumPy U,V,W; every rank "
                       "uploads 1/N of the particle table (all-gathered over NVLink) and its own mask slab and "
                       "returns its own z-slab",
               "host_numa_binding": numa}
        sink = os.path.join(ROOT, "profiles", f"r02_host_sink_n{world}.json")
        if os.path.exists(sink):  # measured floor: all ranks draining their slabs into pinned host memory at once
            try:
                floor_ms = 1e3 * json.load(open(sink))["d2h_sink"]["best_s"]
                e2e["host_sink_floor_ms"] = floor_ms
                e2e["frac_of_host_sink"] = floor_ms / e2e_ms
            except Exception:
                pass
        # the unmodified reference call (main.py:184-192 passes no mask): every voxel is interpolated
        if not args.no_all_voxel:
            barrier()
            t0 = time.perf_counter()
            plugin_call(None)
            torch.cuda.synchronize()
            t3 = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            all_voxel_ms = float(t3.item())
        gi._dev_results.clear()
        torch.cuda.empty_cache()

        # ---- engine-level e2e (round-1 definition): pinned buffers allocated once, particles enter through
        #      rank 0 and are broadcast over NVLink, U,V,W drained chunk by chunk
        out = torch.empty((3, nzl, n, n), dtype=torch.float32, device=dev)
        hp = torch.empty(points.shape, dtype=torch.float64, pin_memory=True).copy_(points)
        hv = torch.empty(values.shape, dtype=torch.float64, pin_memory=True).copy_(values)
        hm = torch.empty(mask_slab.shape, dtype=torch.uint8, pin_memory=True).copy_(mask_slab)
        hout = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)
        dp, dv, dm = (torch.empty_like(t, device=dev) for t in (hp, hv, hm))

        def engine_step():
            if rank == 0 or world == 1:
                dp.copy_(hp, non_blocking=True)
                dv.copy_(hv, non_blocking=True)
            if world > 1:
                dist.broadcast(dp, 0)
                dist.broadcast(dv, 0)
            dm.copy_(hm, non_blocking=True)
            eng.build(dp, dv)
            _, fin = eng.interpolate_to_host(ax, ax, ax[z0:z1], hout, mask=dm, dev_out=out, method=method, k=k)
            torch.cuda.current_stream().wait_event(fin)

        engine_step()
        barrier()
        a, b = ev(), ev()
        a.record()
        for _ in range(n_e2e):
            engine_step()
        b.record()
        barrier()
        t4 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t4, op=dist.ReduceOp.MAX)
        eng_ms = float(t4.item()) / n_e2e
        e2e_engine = {"value": total_pore / (eng_ms * 1e-3), "unit": UNIT, "ms_per_step": eng_ms,
                      "via": "PTVEngine.interpolate_to_host (pinned buffers allocated once)"}
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
    used_stream = kst["used_stream"]
    prof = knn_profile(args.workload, world) if used_stream else None
    roofline = {"kernel": "knn_duo_kernel (+ knn_interp_kernel on handed-over tiles)" if used_stream
                else "knn_interp_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture of THIS
                # kernel source on this workload (null when the source changed since the capture)
                "traffic": prof["dram_bytes_per_launch"] if prof else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": knn_ms,
                "note": "kNN selection is SM-issue bound, not HBM bound (DESIGN.md); the HBM-bound kernels are "
                        "listed under roofline_other"}
    # What actually bounds that kernel: warp-instruction issue.  Instructions per launch come from the same
    # capture (smsp__inst_executed.sum); the time is this run's; peak = 148 SMs x 4 schedulers x one warp
    # instruction per clock at the maximum SM clock.
    roofline_issue = None
    if prof:
        issue_peak = 148 * 4 * 1.965e9
        issue_ach = prof["warp_inst_per_launch"] / (knn_ms * 1e-3)
        roofline_issue = {"kernel": prof["kernel"], "bound": "sm_issue", "achieved": issue_ach / 1e9,
                          "peak": issue_peak / 1e9, "unit": "G warp-inst/s", "frac": issue_ach / issue_peak,
                          "warp_inst_per_launch": prof["warp_inst_per_launch"],
                          "thread_inst_per_pore_voxel": prof["warp_inst_per_launch"] * prof["threads_per_inst"] / total_pore,
                          "source": os.path.relpath(KNN_PROFILE, ROOT) + " (kernel source hash checked)"}
        if work:
            cands = work["candidates_per_voxel_histogram_pass"] + work["candidates_per_voxel_classify_pass"]
            work["thread_inst_per_candidate"] = roofline_issue["thread_inst_per_pore_voxel"] / max(cands, 1.0)
    st_ms = float(np.mean(phase_ms["stencils"]))
    st_bytes = 17.0 * nzl * n * n  # fused divergence + flux + statistics: 12 B u,v,w + 1 B mask read, 4 B div written
    bd_ms = float(np.mean(phase_ms["build"]))
    roofline_other = [{"kernel": "div_flux_bulk_kernel (1 launch + halo/reduce)", "bound": "hbm",
                       "achieved": st_bytes / (st_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                       "frac": st_bytes / (st_ms * 1e-3) / 1e9 / peak, "ms": st_ms},
                      {"kernel": "hash build", "bound": "hbm",
                       "achieved": 112.0 * npart / (bd_ms * 1e-3) / 1e9, "peak": peak,
                       "unit": "GB/s", "frac": 112.0 * npart / (bd_ms * 1e-3) / 1e9 / peak, "ms": bd_ms}]

    cpu_baseline = None
    cpu_rows = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_sample_run(points_np, values_np, mask_sample_np, n, method, k, args.cpu_sample_voxels, workers=1,
                           tree=tree)
        if r["build_s"] is None:  # the parity check built the tree: time a build of its own
            from scipy.spatial import KDTree
            t0 = time.perf_counter()
            KDTree(points_np)
            r["build_s"] = time.perf_counter() - t0
        v, t_total = cpu_throughput(r, n ** 3, total_pore)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                        "host_cores_available": os.cpu_count(),
                        "sample": f"{r['sample_vox']} voxels (3 z-plane windows) of the same workload through the "
                                  f"oracle port as the reference runs it (cKDTree.query workers=1 + NumPy weights): "
                                  f"{r['query_s']:.1f}s query + {r['build_s']:.1f}s tree build; scaled to {n**3} "
                                  f"voxels -> {t_total:.0f}s per frame"}
        if not args.no_cpu_rows:
            cpu_rows = {}
            rb = cpu_sample_run(points_np, values_np, mask_sample_np, n, method, k, args.cpu_sample_voxels, workers=-1,
                                tree=r["tree"], build_s=r["build_s"])
            vb, _ = cpu_throughput(rb, n ** 3, total_pore)
            cpu_rows["workers_all"] = {"value": vb, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                       "mode": "cKDTree.query(workers=-1): not reference behaviour, best case"}
            # the reference's own parallel mode exists for RBF only: CPU pool vs the GPU RBF kernel, same windows
            wins = [(z, (n - 8) // 2, 8) for z in (n // 6, n // 2)]
            npz = os.path.join(tempfile.gettempdir(), f"ptv_bench_rbf_{os.getpid()}.npz")
            np.savez(npz, points=points_np, values=values_np, n=n, total_vox=n ** 3, total_pore=total_pore,
                     windows=np.array(wins))
            try:
                res = subprocess.run([sys.executable, os.path.abspath(__file__), "--rbf-rows-from", npz],
                                     capture_output=True, text=True, timeout=900)
                cpu_rows["rbf_process_pool"] = json.loads(res.stdout.strip().splitlines()[-1])
            except Exception as exc:  # keep the bench line even if the pool row fails
                cpu_rows["rbf_process_pool"] = {"error": repr(exc)[:200]}
            finally:
                try:
                    os.unlink(npz)
                except OSError:
                    pass
            # GPU local RBF (k=20) on whole planes of the same workload
            eng.build(points, values)
            zs = [n // 6, n // 2]
            rbf_out = torch.empty((3, 1, n, n), dtype=torch.float32, device=dev)
            eng.interpolate(ax, ax, ax[zs[0]:zs[0] + 1], method="rbf", k=20, out=rbf_out)
            torch.cuda.synchronize()
            a, b = ev(), ev()
            a.record()
            for z in zs:
                eng.interpolate(ax, ax, ax[z:z + 1], method="rbf", k=20, out=rbf_out)
            b.record()
            torch.cuda.synchronize()
            cpu_rows["rbf_gpu"] = {"voxels_per_sec_all": len(zs) * n * n / (a.elapsed_time(b) * 1e-3),
                                   "mode": "method='rbf', rbf_neighbors=20 on the CUDA path, two whole z-planes, "
                                           "every voxel (no mask)"}

    config = core_config(args.workload, n, npart, method, k, total_pore)
    details = {"porosity": total_pore / n ** 3, "parallelism": f"z-slab x{world}, cuts balanced by pore voxels per plane", "slab_planes_rank0": nzl,
               "l2_policy": "inputs_larger_than_L2" if nzl * n * n * 13 > 126e6 else "small_workload_fits_L2",
               "mask_skip": True, "all_voxels_per_sec": n ** 3 / (ms_per_step * 1e-3),
               "all_voxel_ms": all_voxel_ms,
               "phase_ms_rank0": {p: float(np.mean(v)) for p, v in phase_ms.items()},
               "mean_abs_div": mean_abs_div, "knn_work": work}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 distances/weights, f32 output", "data": "synthetic",
        "config": config, "details": details,
        "roofline": roofline, "roofline_issue": roofline_issue, "roofline_other": roofline_other,
        "cpu_baseline": cpu_baseline, "cpu_rows": cpu_rows, "e2e": e2e, "e2e_engine": e2e_engine,
        "parity_check": parity, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_c5(args, dev, world, rank, local, lib):
    """Config 5: F PTV frames at 512^3 interpolated back to back, spatial hash rebuilt per frame, in the two
    forms SURVEY.md 8(e) asks for -- frame-parallel (one whole frame per GPU at a time, no communication;
    the headline `value`) and slab-parallel (every frame sharded over all GPUs; lower latency per frame)."""
    import torch
    import torch.distributed as dist
    from ptv_interpolation_b200 import synthetic
    from ptv_interpolation_b200.distributed import SlabComm, slab_range
    from ptv_interpolation_b200.engine import PTVEngine

    base = synthetic.make_config("c3", device=dev)
    n, method, k = base["n"], base["method"], base["k"]
    mask = base["mask"]
    mask_u8 = mask.view(torch.uint8)
    pore = int(mask.sum())
    F = args.c5_frames
    walls = synthetic.wall_particles(mask, 50, 2)
    # this rank's frames (frame f <-> seed f); every frame has its own particle cloud, the walls are the mask's
    mine = list(range(rank, F, world))
    frames = {}
    for f in (range(F) if world > 1 else mine):
        pts = synthetic.sample_pore_particles(mask, 5_000_000, seed=1000 + f)
        vals = synthetic.sphere_pack_flow(pts, n)
        frames[f] = (torch.cat([pts, walls], 0), torch.cat([vals, torch.zeros_like(walls)], 0))
    npart = frames[mine[0] if mine else 0][0].shape[0]
    ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)
    eng = PTVEngine(dev)
    out = torch.empty((3, n, n, n), dtype=torch.float32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def sweep_frame_parallel():
        for f in mine:
            p, v = frames[f]
            eng.build(p, v)
            eng.interpolate(ax, ax, ax, mask=mask_u8, method=method, k=k, out=out)

    z0, z1 = slab_range(n, world, rank)

    def sweep_slab_parallel():
        for f in range(F):
            p, v = frames[f]
            eng.build(p, v)
            eng.interpolate(ax, ax, ax[z0:z1], mask=mask_u8[z0:z1], method=method, k=k, out=out[:, z0:z1])

    results = {}
    launches = 0
    clocks = None
    for name, fn in (("frame_parallel", sweep_frame_parallel), ("slab_parallel", sweep_slab_parallel)):
        if name == "slab_parallel" and world == 1:
            continue
        for _ in range(max(1, min(args.warmup, 1))):
            fn()
        sampler = ClockSampler(local) if (rank == 0 and name == "frame_parallel") else None
        barrier()
        if sampler:
            sampler.start()
        l0 = int(lib.ptv_launch_count())
        a, b = ev(), ev()
        a.record()
        for _ in range(args.steps):
            fn()
        b.record()
        barrier()
        if name == "frame_parallel":
            launches = int(lib.ptv_launch_count()) - l0
        if sampler:
            clocks = sampler.stop()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / args.steps
        results[name] = {"ms_per_sweep": ms, "frames_per_sec": F / (ms * 1e-3), "value": F * pore / (ms * 1e-3),
                         "ms_per_frame_latency": ms / F if name == "slab_parallel" else ms / max(len(mine), 1)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peak, peak_src = measured_peak()
    fp = results["frame_parallel"]
    algo = (ALGO_BYTES_PER_VOXEL * n ** 3 + ALGO_BYTES_PER_PARTICLE * npart) * len(mine)
    line = {
        "metric": METRIC, "value": fp["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": fp["ms_per_sweep"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64 distances/weights, f32 output", "data": "synthetic",
        "config": core_config("c5", n, npart, method, k, pore * F),
        "details": {"frames": F, "modes": results, "step": "one sweep over all frames", "headline_mode": "frame_parallel",
                    "l2_policy": "inputs_larger_than_L2"},
        "roofline": {"kernel": "knn_duo_kernel (sibson) + hash build per frame", "bound": "hbm",
                     "achieved": algo / (fp["ms_per_sweep"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": algo / (fp["ms_per_sweep"] * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src},
        "cpu_baseline": None, "e2e": None, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.rbf_rows_from:
        return rbf_pool_row(args.rbf_rows_from)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
