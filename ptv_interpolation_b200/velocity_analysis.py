"""Drop-in mirror of the gradient-stencil part of the reference's ``velocity_analysis.py`` (SURVEY.md 8f
row N3): shear-rate magnitude, vorticity magnitude and viscous dissipation of the interpolated field."""
from __future__ import annotations

import numpy as np

from .engine import default_engine
from .physics import _field_dtype, _to_dev

__all__ = ["compute_strain_rate", "compute_vorticity", "compute_viscous_dissipation"]


def _mask_dev(mask, eng):
    import torch
    return None if mask is None else _to_dev(np.asarray(mask) != 0, eng).view(torch.uint8)


def compute_strain_rate(u, v, w, dx, dy, dz, mask=None, device=None):
    """velocity_analysis.py:10-63."""
    eng = default_engine(device)
    dt = _field_dtype(u, v, w)
    ud, vd, wd = (_to_dev(a, eng, dt) for a in (u, v, w))
    s, _ = eng.strain_vorticity(ud, vd, wd, dx, dy, dz, mask=_mask_dev(mask, eng), vorticity=False)
    return s.cpu().numpy()


def compute_vorticity(u, v, w, dx, dy, dz, mask=None, device=None):
    """velocity_analysis.py:94-120."""
    eng = default_engine(device)
    dt = _field_dtype(u, v, w)
    ud, vd, wd = (_to_dev(a, eng, dt) for a in (u, v, w))
    _, o = eng.strain_vorticity(ud, vd, wd, dx, dy, dz, mask=_mask_dev(mask, eng), strain=False)
    return o.cpu().numpy()


def compute_viscous_dissipation(strain_rate, viscosity, dx=1.0, dy=1.0, dz=1.0, mask=None):
    """velocity_analysis.py:65-92 -- elementwise mu * gamma_dot^2 on the host array it is given."""
    dissipation = viscosity * strain_rate**2
    if mask is not None:
        dissipation[~mask] = 0.0
    return dissipation
