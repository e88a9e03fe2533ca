"""Seeded synthetic inputs for the benchmark configs of BASELINE.json: pore masks
and PTV clouds in VOXEL units, mask True/non-zero = pore, arrays (nz, ny, nx).

Geometry follows the reference generators (generate_sphere_pack.py:25-43 six-sphere hexagonal
pack; generate_cylinders.py:45-49 doublet flow past cylinders); the dense packs used for the
512^3 / 1024^3 configs are FCC lattices of overlapping-free spheres at porosity ~0.4 standing in
for the porous-glass volumes, which are not in the reference tree.  Particle coordinates are
drawn in float32 and returned as float64 (fp32-representable), so the CPU reference and the CUDA
path see identical inputs.  Everything runs on whatever torch device is asked for (CPU for the
small parity cases, the GPU for 512^3+; this is input synthesis, not the product path).
"""
from __future__ import annotations

import math

import torch

__all__ = ["hex6_sphere_pack_mask", "fcc_sphere_pack_mask", "cylinder_array_mask", "sample_pore_particles",
           "sphere_pack_flow", "cylinder_flow", "wall_particles", "make_config"]


def _axes(n, device):
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    z = torch.arange(nz, device=device, dtype=torch.float32)[:, None, None]
    y = torch.arange(ny, device=device, dtype=torch.float32)[None, :, None]
    x = torch.arange(nx, device=device, dtype=torch.float32)[None, None, :]
    return x, y, z, (nx, ny, nz)


def hex6_sphere_pack_mask(n=128, device="cpu"):
    """Six touching spheres, two stacked triangles (generate_sphere_pack.py:25-32), domain padded by
    0.2 (:36-43), voxel i <-> physical min + i*(max-min)/(n-1) per axis (:98-100)."""
    R, D = 0.5, 1.0
    c3y = math.sqrt(3.0) * D / 2.0
    centers = [(0.0, 0.0, 0.0), (D, 0.0, 0.0), (D / 2, c3y, 0.0), (0.0, 0.0, D), (D, 0.0, D), (D / 2, c3y, D)]
    lo = (-R - 0.2, -R - 0.2, -R - 0.2)
    hi = (D + R + 0.2, c3y + R + 0.2, D + R + 0.2)
    x, y, z, (nx, ny, nz) = _axes(n, device)
    px = lo[0] + x * ((hi[0] - lo[0]) / (nx - 1))
    py = lo[1] + y * ((hi[1] - lo[1]) / (ny - 1))
    pz = lo[2] + z * ((hi[2] - lo[2]) / (nz - 1))
    solid = torch.zeros((nz, ny, nx), dtype=torch.bool, device=device)
    for cx, cy, cz in centers:
        solid |= ((px - cx) ** 2 + (py - cy) ** 2 + (pz - cz) ** 2) < R * R
    return ~solid


def fcc_sphere_pack_mask(n, lattice=48.0, porosity=0.40, device="cpu", z0=0, nz_local=None):
    """Dense FCC sphere pack: 4 spheres per cubic cell of edge ``lattice`` voxels, radius set from
    the target porosity (no overlap up to a solid fraction of 0.74).  ``z0``/``nz_local`` build
    only a z-slab of the (n,n,n) volume."""
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    nzl = nz if nz_local is None else nz_local
    a = float(lattice)
    R = a * ((1.0 - porosity) * 3.0 / (16.0 * math.pi)) ** (1.0 / 3.0)
    z = (torch.arange(nzl, device=device, dtype=torch.float32) + float(z0))[:, None, None]
    y = torch.arange(ny, device=device, dtype=torch.float32)[None, :, None]
    x = torch.arange(nx, device=device, dtype=torch.float32)[None, None, :]
    solid = torch.zeros((nzl, ny, nx), dtype=torch.bool, device=device)
    for ox, oy, oz in ((0, 0, 0), (0.5, 0.5, 0), (0.5, 0, 0.5), (0, 0.5, 0.5)):
        dx = x - (torch.round(x / a - ox) + ox) * a
        dy = y - (torch.round(y / a - oy) + oy) * a
        dz = z - (torch.round(z / a - oz) + oz) * a
        solid |= (dx * dx + dy * dy + dz * dz) < R * R
    return ~solid


def cylinder_array_mask(n, pitch=64.0, radius_frac=0.25, device="cpu"):
    """Staggered array of z-aligned cylinders (the 256^3 stand-in for generate_cylinders.py's two
    cylinders in a thin slab): pitch in voxels, radius = radius_frac * pitch."""
    x, y, z, (nx, ny, nz) = _axes(n, device)
    P, R = float(pitch), float(pitch) * radius_frac
    solid2d = torch.zeros((1, ny, nx), dtype=torch.bool, device=device)
    for ox, oy in ((0.0, 0.0), (0.5, 0.5)):
        dx = x - (torch.round(x / P - ox) + ox) * P
        dy = y - (torch.round(y / P - oy) + oy) * P
        solid2d |= (dx * dx + dy * dy) < R * R
    return (~solid2d).expand(nz, ny, nx).contiguous()


def sample_pore_particles(mask, n_points, seed=0, batch=1 << 22):
    """Uniform particles in the pore space of ``mask`` (rejection sampling on the nearest voxel),
    float32 coordinates returned as float64 (Np,3) rows (x,y,z)."""
    device = mask.device
    nz, ny, nx = mask.shape
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    chunks, have = [], 0
    hi = torch.tensor([nx - 1, ny - 1, nz - 1], device=device, dtype=torch.float32)
    while have < n_points:
        p = torch.rand((batch, 3), generator=gen, device=device, dtype=torch.float32) * hi
        ix = torch.round(p[:, 0]).long().clamp_(0, nx - 1)
        iy = torch.round(p[:, 1]).long().clamp_(0, ny - 1)
        iz = torch.round(p[:, 2]).long().clamp_(0, nz - 1)
        keep = mask[iz, iy, ix]
        p = p[keep]
        chunks.append(p)
        have += p.shape[0]
    return torch.cat(chunks, 0)[:n_points].to(torch.float64)


def sphere_pack_flow(points, n):
    """Mean flow along z (generate_sphere_pack.py:86-88: u=v=0, w=1) plus a smooth O(0.3)
    perturbation so interpolation errors are visible; (Np,3) float64 rows (u,v,w)."""
    L = float(n if isinstance(n, int) else max(n))
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    k = 2.0 * math.pi / L
    u = 0.3 * torch.sin(3 * k * y) * torch.cos(2 * k * z)
    v = 0.3 * torch.sin(2 * k * z) * torch.cos(3 * k * x)
    w = 1.0 + 0.3 * torch.sin(2 * k * x) * torch.cos(2 * k * y)
    return torch.stack([u, v, w], -1).to(torch.float32).to(torch.float64)


def cylinder_flow(points, pitch=64.0, radius_frac=0.25, U0=1.0):
    """Uniform flow + doublet of the nearest cylinder (generate_cylinders.py:45-49), w = 0."""
    P, R = float(pitch), float(pitch) * radius_frac
    x, y = points[:, 0], points[:, 1]
    best = None
    for ox, oy in ((0.0, 0.0), (0.5, 0.5)):
        dx = x - (torch.round(x / P - ox) + ox) * P
        dy = y - (torch.round(y / P - oy) + oy) * P
        r2 = dx * dx + dy * dy
        if best is None:
            best = (r2, dx, dy)
        else:
            sel = r2 < best[0]
            best = (torch.where(sel, r2, best[0]), torch.where(sel, dx, best[1]), torch.where(sel, dy, best[2]))
    r2, dx, dy = best
    r2 = r2.clamp_min(1e-6)
    theta = torch.atan2(dy, dx)
    u = U0 * (1 - (R * R / r2) * torch.cos(2 * theta))
    v = -U0 * (R * R / r2) * torch.sin(2 * theta)
    w = torch.zeros_like(u)
    return torch.stack([u, v, w], -1).to(torch.float32).to(torch.float64)


def wall_particles(mask, sampling_step=50, thickness=2):
    """Zero-velocity ghost particles on the grain surfaces, as interpolate_porous_glass.py:68-71 asks for
    (--boundary-particles, sampling 50, thickness 2) and main.py:164-178 builds them: solid voxels within
    ``thickness`` 6-connected dilation steps of the fluid (interpolator.py:256-262), C order, every
    ``sampling_step``-th (:271-274), at their voxel coordinates (bounds (0,n): physical = index, :280-282).
    Input synthesis in torch -- the product path is extract_boundary_particles."""
    fluid = mask.clone()
    for _ in range(int(thickness)):
        d = fluid.clone()
        d[1:] |= fluid[:-1]
        d[:-1] |= fluid[1:]
        d[:, 1:] |= fluid[:, :-1]
        d[:, :-1] |= fluid[:, 1:]
        d[:, :, 1:] |= fluid[:, :, :-1]
        d[:, :, :-1] |= fluid[:, :, 1:]
        fluid = d
    lin = torch.nonzero((fluid & ~mask).reshape(-1)).squeeze(1)[::int(sampling_step)]
    nz, ny, nx = mask.shape
    z, rem = lin // (ny * nx), lin % (ny * nx)
    y, x = rem // nx, rem % nx
    return torch.stack([x, y, z], -1).to(torch.float64)


def make_config(name, device="cpu", seed=None):
    """Inputs of one BASELINE.json config: dict(mask, points, values, n, method, k, ...).
    C1 hex6 128^3/100k, C2 cylinders 256^3/1M, C3 fcc 512^3/5M + wall particles (every 50th boundary voxel,
    thickness 2: the porous-glass launcher's settings), C4 fcc 1024^3/10M."""
    spec = {
        "c1": dict(n=128, npts=100_000, geom="hex6", method="idw", k=50, seed=1),
        "c2": dict(n=256, npts=1_000_000, geom="cyl", method="idw", k=50, seed=2),
        "c3": dict(n=512, npts=5_000_000, geom="fcc", method="sibson", k=50, seed=3, walls=(50, 2)),
        "c4": dict(n=1024, npts=10_000_000, geom="fcc", method="idw", k=50, seed=4),
    }[name]
    n = spec["n"]
    sd = spec["seed"] if seed is None else seed
    if spec["geom"] == "hex6":
        mask = hex6_sphere_pack_mask(n, device)
    elif spec["geom"] == "cyl":
        mask = cylinder_array_mask(n, device=device)
    else:
        mask = fcc_sphere_pack_mask(n, device=device)
    pts = sample_pore_particles(mask, spec["npts"], seed=sd)
    vals = cylinder_flow(pts) if spec["geom"] == "cyl" else sphere_pack_flow(pts, n)
    if spec.get("walls"):
        wp = wall_particles(mask, *spec["walls"])
        pts = torch.cat([pts, wp], 0)
        vals = torch.cat([vals, torch.zeros_like(wp)], 0)
    out = dict(spec)
    out.update(mask=mask, points=pts, values=vals, bounds=((0, n), (0, n), (0, n)))
    return out
