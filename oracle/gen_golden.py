#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run from the repo root:  python oracle/gen_golden.py
Needs /root/reference (read-only mount, absent on the GPU box) -- the committed .npz files
are what travels.  ``tifffile`` is not installed here, so an empty stub module is placed in
``sys.modules`` before importing ``interpolator`` (only ``load_mask`` uses it).
All inputs are seeded; particle coordinates are fp32-representable float64.
"""
import os
import sys
import types

import numpy as np
import pandas as pd

REF = os.environ.get("PTV_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference():
    sys.modules.setdefault("tifffile", types.ModuleType("tifffile"))
    sys.path.insert(0, REF)
    import interpolator as ref_interp  # noqa
    import physics as ref_phys  # noqa
    sys.path.pop(0)
    return ref_interp, ref_phys


def f32r(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def make_df(points, values):
    return pd.DataFrame({"x": points[:, 0], "y": points[:, 1], "z": points[:, 2],
                         "u": values[:, 0], "v": values[:, 1], "w": values[:, 2]})


def main():
    ri, rp = load_reference()
    os.makedirs(OUT, exist_ok=True)
    from scipy.spatial import KDTree

    # ---- case A: random cloud, anisotropic grid, idw / sibson / rbf / nearest
    rng = np.random.default_rng(101)
    bounds = ((0, 14), (0, 12), (0, 10))
    res = (13, 11, 9)
    n = 600
    pts = f32r(rng.uniform([-1, -1, -1], [15, 13, 11], size=(n, 3)))
    vals = f32r(rng.normal(size=(n, 3)))
    df = make_df(pts, vals)
    grid, axes = ri.create_grid(bounds, res)
    fc = np.stack([grid[0].ravel(), grid[1].ravel(), grid[2].ravel()], -1)
    tree = KDTree(pts)
    out = {"points": pts, "values": vals, "bounds": np.array(bounds, dtype=np.float64),
           "res": np.array(res), "ax_x": axes[0], "ax_y": axes[1], "ax_z": axes[2]}
    for k in (1, 8, 50):
        d, i = tree.query(fc, k=k)
        if k == 1:
            d, i = d[:, None], i[:, None]
        out[f"knn_d_k{k}"] = d
        out[f"knn_i_k{k}"] = i
    for name, kw in (("idw_k50", dict(method="idw")),
                     ("idw_k8_p3", dict(method="idw", idw_neighbors=8, idw_power=3.0)),
                     ("idw_k8_p15", dict(method="idw", idw_neighbors=8, idw_power=1.5)),
                     ("sibson_k30", dict(method="sibson")),
                     ("sibson_k12", dict(method="sibson", sibson_neighbors=12)),
                     ("rbf_k20", dict(method="rbf")),
                     ("rbf_k12_s01", dict(method="rbf", rbf_neighbors=12, smoothing=0.1)),
                     ("rbf_k40", dict(method="rbf", rbf_neighbors=40)),
                     ("rbf_cubic_k20", dict(method="rbf", rbf_kernel="cubic")),
                     ("rbf_linear_k15_s005", dict(method="rbf", rbf_kernel="linear", rbf_neighbors=15, smoothing=0.05)),
                     ("rbf_quintic_k30", dict(method="rbf", rbf_kernel="quintic", rbf_neighbors=30)),
                     ("rbf_k60_s001", dict(method="rbf", rbf_neighbors=60, smoothing=0.01)),
                     ("nearest", dict(method="nearest"))):
        U, V, W = ri.interpolate_field(df, grid, **kw)
        out[name] = np.stack([U, V, W], 0)
    np.savez_compressed(os.path.join(OUT, "case_a_interp.npz"), **out)

    # ---- case B: lattice (tie-heavy) boundary particles + random cloud, idw
    rng = np.random.default_rng(202)
    nzm, nym, nxm = 12, 12, 12
    zz, yy, xx = np.meshgrid(np.arange(nzm), np.arange(nym), np.arange(nxm), indexing="ij")
    mask_raw = ((xx - 5.5) ** 2 + (yy - 5.5) ** 2 + (zz - 5.5) ** 2) > 3.6 ** 2  # True = fluid
    bounds_b = ((0, nxm), (0, nym), (0, nzm))
    outb = {"mask_raw": mask_raw, "bounds": np.array(bounds_b, dtype=np.float64)}
    for th, st in ((1, 1), (2, 1), (2, 3), (3, 5)):
        bx, by, bz = ri.extract_boundary_particles(mask_raw, bounds_b, sampling_step=st, thickness=th)
        outb[f"bp_t{th}_s{st}"] = np.stack([bx, by, bz], 0).astype(np.float64)
    bx, by, bz = ri.extract_boundary_particles(mask_raw, bounds_b, sampling_step=1, thickness=1)
    nfl = 500
    p = f32r(rng.uniform(0, 11, size=(4 * nfl, 3)))
    keep = ((p[:, 0] - 5.5) ** 2 + (p[:, 1] - 5.5) ** 2 + (p[:, 2] - 5.5) ** 2) > 3.6 ** 2
    p = p[keep][:nfl]
    v = f32r(np.stack([-(p[:, 1] - 5.5), p[:, 0] - 5.5, 1.0 + 0.1 * p[:, 2]], -1))
    pts_b = np.concatenate([p, np.stack([bx, by, bz], -1)], 0)
    vals_b = np.concatenate([v, np.zeros((len(bx), 3))], 0)
    gridb, axb = ri.create_grid(bounds_b, 12)
    U, V, W = ri.interpolate_field(make_df(pts_b, vals_b), gridb, method="idw", idw_neighbors=20)
    outb.update(points=pts_b, values=vals_b, idw_k20=np.stack([U, V, W], 0))
    msk = ri.sample_mask_on_grid(mask_raw, gridb, bounds_b)
    outb["mask_grid"] = msk
    np.savez_compressed(os.path.join(OUT, "case_b_boundary.npz"), **outb)

    # ---- case C: mask resampling (downscale, crop-like bounds, out of bounds, 0.5 ties)
    rng = np.random.default_rng(303)
    outc = {}
    mraw = rng.random((20, 18, 16)) > 0.45
    outc["mask_raw"] = mraw
    specs = {
        "same": (((0, 16), (0, 18), (0, 20)), ((0, 16), (0, 18), (0, 20)), (16, 18, 20)),
        "down2": (((0, 16), (0, 18), (0, 20)), ((0, 16), (0, 18), (0, 20)), (8, 9, 10)),
        "down3": (((0, 16), (0, 18), (0, 20)), ((0, 16), (0, 18), (0, 20)), (5, 6, 7)),
        "shift": (((3, 19), (-2, 16), (1.5, 21.5)), ((0, 16), (0, 18), (0, 20)), (11, 13, 9)),
        "up": (((0, 16), (0, 18), (0, 20)), ((2, 9), (3, 11), (4, 12)), (15, 17, 19)),
    }
    for name, (braw, bgrid, resg) in specs.items():
        g, _ = ri.create_grid(bgrid, resg)
        outc[f"{name}_out"] = ri.sample_mask_on_grid(mraw, g, braw)
        outc[f"{name}_braw"] = np.array(braw, dtype=np.float64)
        outc[f"{name}_bgrid"] = np.array(bgrid, dtype=np.float64)
        outc[f"{name}_res"] = np.array(resg)
    np.savez_compressed(os.path.join(OUT, "case_c_mask.npz"), **outc)

    # ---- case D: divergence stencil + flux + statistics
    rng = np.random.default_rng(404)
    shape = (11, 9, 7)
    u, v, w = (rng.normal(size=shape) for _ in range(3))
    m = rng.random(shape) > 0.35
    dx, dy, dz = 1.25, 0.75, 2.0
    div = rp.compute_consistent_divergence(u, v, w, m, dx, dy, dz)
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    sys.path.insert(0, REF)
    import plot_flux as pf
    sys.path.pop(0)
    np.savez_compressed(os.path.join(OUT, "case_d_divergence.npz"), u=u, v=v, w=w, mask=m,
                        h=np.array([dx, dy, dz]), div=div,
                        q_xy=pf.calculate_flux_xy(w, dx, dy), q_xz=pf.calculate_flux_xz(v, dx, dz),
                        q_yz=pf.calculate_flux_yz(u, dy, dz),
                        mid_x=np.sum(u[:, :, shape[2] // 2]) * dy * dz,
                        mean_abs_div=np.mean(np.abs(div[m])))

    # ---- case E: the reference's own test (test_parallel.py:6-27), rbf n_jobs=2
    dfe = pd.DataFrame({"x": [0, 10, 0, 10, 5], "y": [0, 0, 10, 10, 5], "z": [0, 0, 0, 0, 5],
                        "u": [1, 1, 1, 1, 2], "v": [0, 0, 0, 0, 0], "w": [0, 0, 0, 0, 0]})
    ge, _ = ri.create_grid(((0, 10), (0, 10), (0, 10)), 10)
    U, V, W = ri.interpolate_field(dfe, ge, method="rbf", n_jobs=2)
    np.savez_compressed(os.path.join(OUT, "case_e_test_parallel.npz"), uvw=np.stack([U, V, W], 0))
    # ---- case F: kNN median/MAD outlier filter (filtering.py:5-58), with planted outliers
    sys.path.insert(0, REF)
    import filtering as rf
    sys.path.pop(0)
    rng = np.random.default_rng(606)
    pf_ = f32r(rng.uniform(0, 20, size=(3000, 3)))
    vf_ = f32r(np.stack([1 + 0.1 * pf_[:, 1], 0.2 * np.sin(pf_[:, 0]), 0.05 * pf_[:, 2]], -1)
               + 0.02 * rng.normal(size=(3000, 3)))
    bad = rng.choice(3000, 60, replace=False)
    vf_[bad] *= rng.uniform(3, 8, size=(60, 1))
    dff = make_df(pf_, vf_)
    outf = {"points": pf_, "values": vf_}
    for kf, thr in ((25, 3.0), (10, 2.0), (24, 3.5)):
        kept = rf.remove_outliers_knn(dff.copy(), k=kf, threshold=thr)
        outf[f"kept_k{kf}_t{thr}"] = kept[["x", "y", "z", "u", "v", "w"]].values
    np.savez_compressed(os.path.join(OUT, "case_f_filter.npz"), **outf)
    # ---- case G: strain rate / vorticity / dissipation (velocity_analysis.py:10-120)
    sys.path.insert(0, REF)
    import velocity_analysis as va
    sys.path.pop(0)
    rng = np.random.default_rng(707)
    shape = (9, 8, 10)
    ug, vg, wg = (rng.normal(size=shape) for _ in range(3))
    mg = rng.random(shape) > 0.3
    hs = (1.25, 0.75, 2.0)
    sr = va.compute_strain_rate(ug, vg, wg, *hs, mask=mg)
    np.savez_compressed(os.path.join(OUT, "case_g_analysis.npz"), u=ug, v=vg, w=wg, mask=mg, h=np.array(hs),
                        strain=sr, strain_nomask=va.compute_strain_rate(ug, vg, wg, 1.0, 1.0, 1.0),
                        vort=va.compute_vorticity(ug, vg, wg, *hs, mask=mg),
                        diss=va.compute_viscous_dissipation(sr.copy(), 1.3e-3, mask=mg))
    # ---- case H: projection cleaning (physics.py:149-209): Laplacian, lsqr, correction, 3 iterations
    import contextlib, io
    rng = np.random.default_rng(808)
    shape = (12, 11, 10)
    zz, yy, xx = np.meshgrid(*(np.arange(s_, dtype=float) for s_ in shape), indexing="ij")
    mh = ((xx - 4.5) ** 2 + (yy - 5.0) ** 2 + (zz - 6.0) ** 2) > 2.7 ** 2
    mh &= ~((xx > 7) & (yy > 8))
    uh = (1.0 + 0.3 * np.sin(0.5 * yy) + 0.05 * rng.normal(size=shape)) * mh
    vh = (0.2 * np.cos(0.4 * xx) + 0.05 * rng.normal(size=shape)) * mh
    wh = (0.1 * np.sin(0.3 * zz + 0.2 * xx) + 0.05 * rng.normal(size=shape)) * mh
    hh = (1.0, 1.5, 0.8)
    with contextlib.redirect_stdout(io.StringIO()):
        uc, vc, wc = rp.clean_divergence_projection(uh, vh, wh, mh, *hh, iterations=3)
        u1, v1, w1 = rp.clean_divergence_projection(uh, vh, wh, mh, *hh, iterations=1)
    A, idx_map = rp.build_laplacian_matrix(mh, *hh)
    xt = rng.normal(size=A.shape[0])
    np.savez_compressed(os.path.join(OUT, "case_h_projection.npz"), u=uh, v=vh, w=wh, mask=mh, h=np.array(hh),
                        u3=uc, v3=vc, w3=wc, u1=u1, v1=v1, w1=w1, lap_x=xt, lap_Ax=A @ xt,
                        div0=rp.compute_consistent_divergence(uh, vh, wh, mh, *hh),
                        div3=rp.compute_consistent_divergence(uc, vc, wc, mh, *hh))
    gen_linear(ri)
    print("golden vectors written to", os.path.normpath(OUT))


def _simplex_rows(pts, fc):
    """Vertex rows (ascending) of the Delaunay tetrahedron scipy finds for each query, -1 outside the hull."""
    from scipy.spatial import Delaunay
    tri = Delaunay(pts)
    s = tri.find_simplex(fc)
    rows = np.where(s[:, None] >= 0, np.sort(tri.simplices[np.maximum(s, 0)], axis=1), -1)
    return rows.astype(np.int64)


def gen_linear(ri):
    """case I: method='linear' (interpolator.py:197 -> griddata -> Qhull Delaunay), the reference's default.
    `python oracle/gen_golden.py linear` writes only this file."""
    out = {}
    # I1: case A's cloud and grid (all voxels well inside the hull)
    rng = np.random.default_rng(101)
    bounds = ((0, 14), (0, 12), (0, 10))
    res = (13, 11, 9)
    pts = f32r(rng.uniform([-1, -1, -1], [15, 13, 11], size=(600, 3)))
    vals = f32r(rng.normal(size=(600, 3)))
    grid, _ = ri.create_grid(bounds, res)
    fc = np.stack([grid[0].ravel(), grid[1].ravel(), grid[2].ravel()], -1)
    U, V, W = ri.interpolate_field(make_df(pts, vals), grid, method="linear")
    out.update(a_points=pts, a_values=vals, a_bounds=np.array(bounds, dtype=np.float64), a_res=np.array(res),
               a_uvw=np.stack([U, V, W], 0), a_simplex=_simplex_rows(pts, fc))
    # I2: a cloud smaller than the grid: a third of the voxels lie outside the convex hull (fill_value 0)
    rng = np.random.default_rng(909)
    pts = f32r(rng.uniform(2.5, 9.5, size=(400, 3)))
    vals = f32r(np.stack([1.0 + 0.2 * pts[:, 1], np.sin(pts[:, 0]), 0.1 * pts[:, 2] ** 2], -1))
    bounds = ((0, 13), (0, 13), (0, 13))
    grid, _ = ri.create_grid(bounds, (14, 12, 10))
    fc = np.stack([grid[0].ravel(), grid[1].ravel(), grid[2].ravel()], -1)
    U, V, W = ri.interpolate_field(make_df(pts, vals), grid, method="linear")
    out.update(b_points=pts, b_values=vals, b_bounds=np.array(bounds, dtype=np.float64), b_res=np.array((14, 12, 10)),
               b_uvw=np.stack([U, V, W], 0), b_simplex=_simplex_rows(pts, fc))
    # I3: case B's cloud: random pore particles + zero-velocity wall particles on the voxel lattice
    # (co-spherical by construction; the grid points coincide with lattice sites) -- values only
    cb = np.load(os.path.join(OUT, "case_b_boundary.npz"))
    gridb, _ = ri.create_grid(((0, 12), (0, 12), (0, 12)), 12)
    U, V, W = ri.interpolate_field(make_df(cb["points"], cb["values"]), gridb, method="linear")
    out.update(c_uvw=np.stack([U, V, W], 0))
    np.savez_compressed(os.path.join(OUT, "case_i_linear.npz"), **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["linear"]:
        gen_linear(load_reference()[0])
    else:
        main()
