"""The hot path end to end on device tensors (main.py steps 4-6 + the stencils that consume the
grid): hash build -> fused kNN/weights/mask -> divergence (+ halos) -> flux profiles and
mean|div| (+ reductions).  Used by bench.py, smoke() and the multi-rank tests."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .distributed import SlabComm
from .engine import PTVEngine


@dataclass
class StepResult:
    uvw: torch.Tensor            # (3, nz_local, ny, nx)
    div: torch.Tensor            # (nz_local, ny, nx)
    q_xy: torch.Tensor           # (nz,) global, unscaled plane sums of w
    q_xz: torch.Tensor           # (ny,)
    q_yz: torch.Tensor           # (nx,)
    mean_abs_div: torch.Tensor   # 0-d float64
    n_fluid: torch.Tensor        # 0-d float64


def hot_path_step(eng: PTVEngine, points, values, ax_x, ax_y, ax_z, mask_slab, comm: SlabComm, method="idw",
                  k=50, idw_power=2.0, spacing=(1.0, 1.0, 1.0), out=None, out_dtype=torch.float32,
                  rebuild=True, mark=None, overlap_halos=False) -> StepResult:
    """One pass over one PTV frame for this rank's z-slab.  ``ax_z`` is the FULL z axis; the slab is
    comm.z0:comm.z1.  ``mask_slab`` is the (nz_local, ny, nx) uint8 pore mask of the slab.  ``mark(label)``
    is called after the hash build ("built") and after the last interpolation launch ("interpolated").
    ``overlap_halos``: interpolate the slab's first and last 32 planes first and let the halo exchange travel while
    the interior is searched.  Off by default: the exchange is 10 MB over NVLink (0.1 ms exposed), while three
    launches instead of one cost three kernel tails -- measured at 8 GPUs 35.0 ms per frame without, 35.8 ms with
    (profiles/r02_v4_bench_c4_n8*.json), at 2 GPUs 136.2 vs 136.7 ms."""
    mark = mark or (lambda label: None)
    slab_hash = rebuild and comm.world > 1 and method in ("idw", "sibson")
    if slab_hash:
        # north_star: "each GPU holds its slab's particles plus a halo" -- bin only z within the slab +- a halo
        # of expected k-neighbour radii; a search that leaves the range is detected below
        key = (ax_z.data_ptr(), comm.z0, comm.z1)
        if getattr(comm, "_zrange_key", None) != key:  # two host reads, once per grid (not per frame)
            comm._zrange = (float(ax_z[comm.z0]), float(ax_z[comm.z1 - 1]))
            comm._zrange_key = key
        eng.build_slab(points, values, min(comm._zrange), max(comm._zrange), k)
    elif rebuild:
        eng.build(points, values)
    mark("built")
    az = ax_z[comm.z0:comm.z1]
    nzl = comm.z1 - comm.z0
    nx, ny = ax_x.numel(), ax_y.numel()
    kw = dict(method=method, k=k, idw_power=idw_power)
    edge = 32  # z-extent of one CTA region of the streaming kernel
    if overlap_halos and comm.world > 1 and nzl >= 3 * edge:
        # the slab's first and last planes feed the neighbours' divergence stencils: interpolate the two
        # boundary chunks first, start the halo exchange, and let it travel while the interior is searched
        if out is None:
            out = torch.empty((3, nzl, ny, nx), dtype=out_dtype, device=eng.device)
        for a, b in ((0, edge), (nzl - edge, nzl)):
            eng.interpolate(ax_x, ax_y, az[a:b], mask=mask_slab[a:b], out=out[:, a:b], **kw)
        pending = comm.post_halos(out[2, 0], out[2, nzl - 1], mask_slab[0])
        eng.interpolate(ax_x, ax_y, az[edge:nzl - edge], mask=mask_slab[edge:nzl - edge], out=out[:, edge:nzl - edge], **kw)
        uvw = out
        mark("interpolated")
        w_below, w_above, m_above = comm.wait_halos(pending)
    else:
        uvw = eng.interpolate(ax_x, ax_y, az, mask=mask_slab, out_dtype=out_dtype, out=out, **kw)
        mark("interpolated")
        w_below, w_above, m_above = comm.exchange_halos(uvw[2], mask_slab)
    dx, dy, dz = spacing
    div, stats, q_xy, q_xz, q_yz, acc = eng.divergence_flux(uvw[0], uvw[1], uvw[2], mask_slab, dx, dy, dz,
                                                            w_below=w_below, w_above=w_above, mask_above=m_above,
                                                            z0=comm.z0, nz_global=comm.nz, extra=1)
    if slab_hash:
        eng.clip_violations_to(acc[-1:])  # rides along in the all-reduce below
    comm.reduce_profiles_(acc)  # one all-reduce: (sum|div|, n_fluid), Q_xy, Q_xz, Q_yz (+ halo violations)
    if slab_hash:
        # exact fall-back: some voxel's k-th neighbour may lie beyond the halo -> redo the frame on the full
        # hash, on every rank (the sum is the same everywhere, so the decision is collective and the halo
        # exchange stays matched)
        if float(acc[-1].item()) > 0.0:
            eng.build(points, values)
            return hot_path_step(eng, points, values, ax_x, ax_y, ax_z, mask_slab, comm, method=method, k=k,
                                 idw_power=idw_power, spacing=spacing, out=uvw, out_dtype=out_dtype, rebuild=False,
                                 overlap_halos=overlap_halos)
    return StepResult(uvw, div, q_xy, q_xz, q_yz, stats[0] / stats[1], stats[1])
