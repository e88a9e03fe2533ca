"""Drop-in mirror of the grid-physics functions that consume the interpolated field:
``compute_consistent_divergence`` (reference physics.py:6-53), the plane fluxes of
plot_flux.py:6-16, the mid-plane X flux (physics.py:160-165) and mean|div| over fluid voxels
(physics.py:174, view_divergence.py:45-46) -- all on the CUDA path, no CPU fallback."""
from __future__ import annotations

import numpy as np

from .engine import default_engine

__all__ = ["compute_consistent_divergence", "calculate_flux_xy", "calculate_flux_xz", "calculate_flux_yz",
           "mid_plane_x_flux", "mean_abs_divergence"]


def _to_dev(a, eng, dtype=None):
    import torch
    a = np.ascontiguousarray(a)
    if dtype is not None:
        a = a.astype(dtype, copy=False)
    return torch.from_numpy(a).to(eng.device)


def _field_dtype(*arrs):
    return np.float32 if all(np.asarray(a).dtype == np.float32 for a in arrs) else np.float64


def compute_consistent_divergence(u, v, w, mask, dx, dy, dz, device=None):
    """physics.py:6-53.  float32 fields stay float32 (arithmetic in float64 inside the kernel);
    anything else is computed and returned in float64, bit-identical to the reference."""
    eng = default_engine(device)
    dt = _field_dtype(u, v, w)
    ud, vd, wd = (_to_dev(a, eng, dt) for a in (u, v, w))
    md = _to_dev(np.asarray(mask) != 0, eng).view(__import__("torch").uint8)
    div = eng.divergence(ud, vd, wd, md, dx, dy, dz)
    return div.cpu().numpy()


def mean_abs_divergence(u, v, w, mask, dx, dy, dz, device=None):
    """mean(|div[mask]|) as printed at physics.py:174-175 / view_divergence.py:45-51."""
    eng = default_engine(device)
    dt = _field_dtype(u, v, w)
    ud, vd, wd = (_to_dev(a, eng, dt) for a in (u, v, w))
    md = _to_dev(np.asarray(mask) != 0, eng).view(__import__("torch").uint8)
    _, stats = eng.divergence(ud, vd, wd, md, dx, dy, dz, with_stats=True)
    s, c = stats.cpu().numpy()
    return float(s / c) if c > 0 else float("nan")


def _profiles(u=None, v=None, w=None, device=None):
    eng = default_engine(device)
    f = [None if a is None else _to_dev(a, eng, _field_dtype(a)) for a in (u, v, w)]
    qxy, qxz, qyz = eng.flux_profiles(*f)
    return qxy.cpu().numpy(), qxz.cpu().numpy(), qyz.cpu().numpy()


def calculate_flux_xy(w_field, dx, dy, device=None):
    """plot_flux.py:6-8 -- sum over (y, x) of W per z-plane, times dx*dy."""
    return _profiles(w=w_field, device=device)[0] * dx * dy


def calculate_flux_xz(v_field, dx, dz, device=None):
    """plot_flux.py:10-12."""
    return _profiles(v=v_field, device=device)[1] * dx * dz


def calculate_flux_yz(u_field, dy, dz, device=None):
    """plot_flux.py:14-16."""
    return _profiles(u=u_field, device=device)[2] * dy * dz


def mid_plane_x_flux(u_field, dy, dz, device=None):
    """physics.py:160-165 -- net flux through the middle YZ plane."""
    nx = np.asarray(u_field).shape[2]
    return float(_profiles(u=u_field, device=device)[2][nx // 2] * dy * dz)
