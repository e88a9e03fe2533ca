"""Drop-in mirror of the grid-physics functions that consume the interpolated field:
``compute_consistent_divergence`` (reference physics.py:6-53), the plane fluxes of
plot_flux.py:6-16, the mid-plane X flux (physics.py:160-165) and mean|div| over fluid voxels
(physics.py:174, view_divergence.py:45-46) -- all on the CUDA path, no CPU fallback."""
from __future__ import annotations

import numpy as np

from .engine import default_engine

__all__ = ["compute_consistent_divergence", "calculate_flux_xy", "calculate_flux_xz", "calculate_flux_yz",
           "mid_plane_x_flux", "mean_abs_divergence", "clean_divergence_projection", "clean_divergence"]


def _to_dev(a, eng, dtype=None):
    """Pageable NumPy array -> device tensor through the cached pinned staging chunks (hostmem.py)."""
    from . import hostmem
    a = np.ascontiguousarray(a)
    if dtype is not None:
        a = a.astype(dtype, copy=False)
    return hostmem.stage_to_device(a, eng.device)


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


def clean_divergence_projection(u, v, w, mask, dx, dy, dz, iterations=3, device=None):
    """physics.py:149-209 -- iterated projection: divergence -> masked Poisson solve (LSQR with SciPy's
    recurrences and stopping rules, on the device) -> staggered-gradient correction.  The whole loop
    runs on device tensors in float64; returns three float64 host arrays like the reference."""
    import torch
    eng = default_engine(device)
    u_c, v_c, w_c = (_to_dev(a, eng, np.float64) for a in (u, v, w))
    md = _to_dev(np.asarray(mask) != 0, eng).view(torch.uint8)
    nx = u_c.shape[2]
    print(f"Starting Iterative Divergence Cleaning ({iterations} iterations)...")

    def report_flux(u_field, label):
        flux = float(u_field[:, :, nx // 2].sum(dtype=torch.float64)) * dy * dz
        print(f"  [{label}] Net X-Flux (mid-plane): {flux:.4e}")

    def mean_abs_div(a, b, c):
        div, st, _, _, _ = eng.divergence_flux(a, b, c, md, dx, dy, dz)
        s, n = st.cpu().numpy()
        return div, (float(s / n) if n > 0 else float("nan"))

    report_flux(u_c, "Initial")
    _, m_div_init = mean_abs_div(u_c, v_c, w_c)
    for i in range(iterations):
        print(f"\n--- Iteration {i+1}/{iterations} ---")
        div, m_div = mean_abs_div(u_c, v_c, w_c)
        print(f"  Current Mean Abs Div: {m_div:.6e}")
        print(f"  Solving Poisson (matrix-free LSQR on {int(md.sum())} fluid voxels)...")
        phi, info = eng.poisson_lsqr(div, md, dx, dy, dz, damp=1e-8, atol=1e-10, btol=1e-10, iter_lim=3000)
        if bool(torch.isnan(phi).any()):
            print("  Warning: Solve failed. Stopping iterations.")
            break
        u_c, v_c, w_c = eng.projection_correct(u_c, v_c, w_c, phi, md, dx, dy, dz)
    _, m_div_final = mean_abs_div(u_c, v_c, w_c)
    print("\n" + "=" * 40)
    print("DIVERGENCE CLEANING COMPLETE")
    print(f"Initial Mean Abs Div: {m_div_init:.6e}")
    print(f"Final Mean Abs Div:   {m_div_final:.6e}")
    print(f"Total Reduction:      {m_div_init/m_div_final:.2f}x")
    report_flux(u_c, "Final")
    print("=" * 40 + "\n")
    return u_c.cpu().numpy(), v_c.cpu().numpy(), w_c.cpu().numpy()


def clean_divergence(u, v, w, mask, dx, dy, dz, iterations=3, method="projection", lambda_reg=1e3, device=None):
    """physics.py:347-354 dispatcher.  The variational method (sparse CG on I + lambda D^T D) is not on
    the CUDA path."""
    if method == "variational":
        raise NotImplementedError("clean_divergence(method='variational') is not implemented on the CUDA path; "
                                  "use method='projection'")
    return clean_divergence_projection(u, v, w, mask, dx, dy, dz, iterations=iterations, device=device)
