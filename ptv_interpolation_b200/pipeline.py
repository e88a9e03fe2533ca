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
                  rebuild=True) -> StepResult:
    """One pass over one PTV frame for this rank's z-slab.  ``ax_z`` is the FULL z axis; the slab is
    comm.z0:comm.z1.  ``mask_slab`` is the (nz_local, ny, nx) uint8 pore mask of the slab."""
    if rebuild:
        eng.build(points, values)
    az = ax_z[comm.z0:comm.z1]
    uvw = eng.interpolate(ax_x, ax_y, az, mask=mask_slab, method=method, k=k, idw_power=idw_power,
                          out_dtype=out_dtype, out=out)
    w_below, w_above, m_above = comm.exchange_halos(uvw[2], mask_slab)
    dx, dy, dz = spacing
    div, stats, q_xy, q_xz, q_yz = eng.divergence_flux(uvw[0], uvw[1], uvw[2], mask_slab, dx, dy, dz, w_below=w_below,
                                                       w_above=w_above, mask_above=m_above)
    comm.reduce_sum_(q_xz, q_yz, stats)
    q_xy = comm.gather_planes(q_xy)
    return StepResult(uvw, div, q_xy, q_xz, q_yz, stats[0] / stats[1], stats[1])
