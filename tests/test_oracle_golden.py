"""Pin the CPU oracle (oracle/reference_port.py) against vectors produced by the live
reference (oracle/gen_golden.py -> tests/golden/*.npz).  fp64 results must be identical."""
import os

import numpy as np
import pytest

from oracle import reference_port as rp


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _grid(bounds, res):
    b = tuple(tuple(float(v) for v in row) for row in bounds)
    return rp.create_grid(b, tuple(int(r) for r in res))


def test_create_grid_axes(golden_dir):
    g = _load(golden_dir, "case_a_interp.npz")
    (X, Y, Z), (x, y, z) = _grid(g["bounds"], g["res"])
    assert np.array_equal(x, g["ax_x"]) and np.array_equal(y, g["ax_y"]) and np.array_equal(z, g["ax_z"])
    assert X.shape == (9, 11, 13)


@pytest.mark.parametrize("k", [1, 8, 50])
def test_knn_matches_ckdtree_and_bruteforce(golden_dir, k):
    g = _load(golden_dir, "case_a_interp.npz")
    grid, _ = _grid(g["bounds"], g["res"])
    fc = rp.flat_coords(grid)
    d, i, d2 = rp.knn_canonical(g["points"], fc, k)
    # random cloud: no ties, so canonical order == cKDTree's own order, bit for bit
    assert np.array_equal(i, g[f"knn_i_k{k}"])
    assert np.array_equal(d, g[f"knn_d_k{k}"])
    db, ib, _ = rp.knn_bruteforce(g["points"], fc, k)
    assert np.array_equal(ib, i) and np.array_equal(db, d)


@pytest.mark.parametrize("name,kw", [
    ("idw_k50", dict(method="idw")),
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
    ("nearest", dict(method="nearest")),
])
def test_interpolate_field_bitexact(golden_dir, name, kw):
    g = _load(golden_dir, "case_a_interp.npz")
    grid, _ = _grid(g["bounds"], g["res"])
    U, V, W = rp.interpolate_field(g["points"], g["values"], grid, chunk_voxels=500, **kw)
    ref = g[name]
    assert np.array_equal(np.stack([U, V, W], 0), ref)


def test_lattice_ties_values_and_boundary_particles(golden_dir):
    g = _load(golden_dir, "case_b_boundary.npz")
    b = tuple(tuple(float(v) for v in row) for row in g["bounds"])
    for th, st in ((1, 1), (2, 1), (2, 3), (3, 5)):
        bx, by, bz = rp.extract_boundary_particles(g["mask_raw"], b, sampling_step=st, thickness=th)
        assert np.array_equal(np.stack([bx, by, bz], 0).astype(np.float64), g[f"bp_t{th}_s{st}"])
    grid, _ = rp.create_grid(b, 12)
    U, V, W = rp.interpolate_field(g["points"], g["values"], grid, method="idw", idw_neighbors=20)
    ref = g["idw_k20"]
    got = np.stack([U, V, W], 0)
    # ties reorder equal-weight terms of the fp64 sums: identical up to summation order
    scale = np.abs(ref).max()
    assert np.max(np.abs(got - ref)) <= 1e-12 * scale
    for use_scipy in (True, False):
        assert np.array_equal(rp.sample_mask_on_grid(g["mask_raw"], grid, b, use_scipy=use_scipy), g["mask_grid"])


@pytest.mark.parametrize("name", ["same", "down2", "down3", "shift", "up"])
def test_mask_sampling(golden_dir, name):
    g = _load(golden_dir, "case_c_mask.npz")
    braw = tuple(tuple(float(v) for v in row) for row in g[f"{name}_braw"])
    bgrid = tuple(tuple(float(v) for v in row) for row in g[f"{name}_bgrid"])
    grid, _ = rp.create_grid(bgrid, tuple(int(r) for r in g[f"{name}_res"]))
    for use_scipy in (True, False):
        out = rp.sample_mask_on_grid(g["mask_raw"], grid, braw, use_scipy=use_scipy)
        assert np.array_equal(out, g[f"{name}_out"]), (name, use_scipy)


def test_divergence_flux_stats(golden_dir):
    g = _load(golden_dir, "case_d_divergence.npz")
    dx, dy, dz = g["h"]
    div = rp.compute_consistent_divergence(g["u"], g["v"], g["w"], g["mask"], dx, dy, dz)
    assert np.array_equal(div, g["div"])
    assert np.array_equal(rp.flux_xy(g["w"], dx, dy), g["q_xy"])
    assert np.array_equal(rp.flux_xz(g["v"], dx, dz), g["q_xz"])
    assert np.array_equal(rp.flux_yz(g["u"], dy, dz), g["q_yz"])
    assert rp.mid_plane_x_flux(g["u"], dy, dz) == g["mid_x"]
    assert rp.mean_abs_div(div, g["mask"]) == g["mean_abs_div"]


def test_known_answers():
    # constant field -> constant; voxel on a particle returns its value
    rng = np.random.default_rng(5)
    pts = rng.uniform(0, 9, size=(200, 3)).astype(np.float32).astype(np.float64)
    pts[0] = (4.0, 4.0, 4.0)
    vals = np.full((200, 3), 7.0)
    grid, _ = rp.create_grid(((0, 10), (0, 10), (0, 10)), 10)
    U, V, W = rp.interpolate_field(pts, vals, grid, method="idw", idw_neighbors=10)
    assert np.allclose(U, 7.0, rtol=1e-12)
    vals = rng.normal(size=(200, 3))
    U, V, W = rp.interpolate_field(pts, vals, grid, method="idw", idw_neighbors=10)
    assert abs(U[4, 4, 4] - vals[0, 0]) <= 1e-7 * max(1.0, abs(vals[0, 0]))
    with pytest.raises(IndexError):
        rp.interpolate_field(pts[:5], vals[:5], grid, method="idw", idw_neighbors=10)


def test_reference_own_test_shape(golden_dir):
    # test_parallel.py:27 only pins the output shape
    g = _load(golden_dir, "case_e_test_parallel.npz")
    assert g["uvw"].shape == (3, 10, 10, 10)
    pts = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [10, 10, 0], [5, 5, 5]], dtype=float)
    vals = np.array([[1, 0, 0], [1, 0, 0], [1, 0, 0], [1, 0, 0], [2, 0, 0]], dtype=float)
    grid, _ = rp.create_grid(((0, 10), (0, 10), (0, 10)), 10)
    U, V, W = rp.interpolate_field(pts, vals, grid, method="rbf")
    assert np.array_equal(np.stack([U, V, W], 0), g["uvw"])


@pytest.mark.parametrize("k,thr", [(25, 3.0), (10, 2.0), (24, 3.5)])
def test_outlier_filter_port(golden_dir, k, thr):
    g = _load(golden_dir, "case_f_filter.npz")
    keep, kth = rp.outlier_keep_mask(g["points"], g["values"], k, thr)
    got = np.concatenate([g["points"], g["values"]], 1)[keep]
    assert np.array_equal(got, g[f"kept_k{k}_t{thr}"])


def test_analysis_stencils_port(golden_dir):
    g = _load(golden_dir, "case_g_analysis.npz")
    dx, dy, dz = g["h"]
    assert np.array_equal(rp.compute_strain_rate(g["u"], g["v"], g["w"], dx, dy, dz, mask=g["mask"]), g["strain"])
    assert np.array_equal(rp.compute_strain_rate(g["u"], g["v"], g["w"], 1.0, 1.0, 1.0), g["strain_nomask"])
    assert np.array_equal(rp.compute_vorticity(g["u"], g["v"], g["w"], dx, dy, dz, mask=g["mask"]), g["vort"])


def test_projection_cleaning_port(golden_dir):
    g = _load(golden_dir, "case_h_projection.npz")
    dx, dy, dz = g["h"]
    A, idx_map = rp.build_laplacian_matrix(g["mask"], dx, dy, dz)
    assert np.array_equal(A @ g["lap_x"], g["lap_Ax"])
    u3, v3, w3 = rp.clean_divergence_projection(g["u"], g["v"], g["w"], g["mask"], dx, dy, dz, iterations=3)
    assert np.array_equal(u3, g["u3"]) and np.array_equal(v3, g["v3"]) and np.array_equal(w3, g["w3"])
    m0 = rp.mean_abs_div(g["div0"], g["mask"])
    m3 = rp.mean_abs_div(g["div3"], g["mask"])
    assert m3 < 0.6 * m0  # the cleaning does reduce the divergence


def test_c_bruteforce_agrees_with_ckdtree_canonical(golden_dir):
    """oracle/knn_brute.c (no SciPy, no NumPy arithmetic) against the cKDTree-based canonical search on a
    random cloud and on the tie-heavy lattice case."""
    from oracle.knn_brute import knn_brute
    rng = np.random.default_rng(12)
    pts = rng.uniform(0, 30, size=(20000, 3)).astype(np.float32).astype(np.float64)
    q = rng.uniform(-2, 32, size=(3000, 3))
    d, i, d2 = rp.knn_canonical(pts, q, 50, workers=-1)
    db, ib, d2b = knn_brute(pts, q, 50)
    assert np.array_equal(ib, i) and np.array_equal(d2b, d2) and np.array_equal(db, d)
    g = _load(golden_dir, "case_b_boundary.npz")
    grid, _ = rp.create_grid(((0, 12), (0, 12), (0, 12)), 12)
    fc = rp.flat_coords(grid)
    d, i, d2 = rp.knn_canonical(g["points"], fc, 20)
    db, ib, _ = knn_brute(g["points"], fc, 20)
    assert np.array_equal(ib, i) and np.array_equal(db, d)
    with pytest.raises(IndexError):
        knn_brute(pts[:5], q, 6)


# ------------------------------------------------------------------ method='linear' (interpolator.py:197)
def test_linear_port_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "case_i_linear.npz"))
    for tag in ("a", "b"):
        grid, _ = rp.create_grid(tuple(map(tuple, g[tag + "_bounds"])), tuple(int(r) for r in g[tag + "_res"]))
        U, V, W = rp.interpolate_field(g[tag + "_points"], g[tag + "_values"], grid, method="linear")
        assert np.array_equal(np.stack([U, V, W]), g[tag + "_uvw"])
        rows, b = rp.delaunay_simplex_rows(g[tag + "_points"], rp.flat_coords(grid))
        assert np.array_equal(rows, g[tag + "_simplex"])
        inside = rows[:, 0] >= 0
        vals = g[tag + "_values"]
        rec = np.einsum("nk,nkc->nc", b[inside], vals[rows[inside]])
        assert np.abs(rec - g[tag + "_uvw"].reshape(3, -1).T[inside]).max() <= 1e-12


def test_linear_programme_restatement_matches_reference(golden_dir):
    """oracle/delaunay_lp.py (no SciPy, no spatial index) finds the tetrahedra the reference evaluated in,
    inside and outside the hull, and reproduces the reference on lattice (co-spherical) wall particles."""
    from oracle.delaunay_lp import linear_interpolate
    g = np.load(os.path.join(golden_dir, "case_i_linear.npz"))
    rng = np.random.default_rng(0)
    for tag in ("a", "b"):
        grid, _ = rp.create_grid(tuple(map(tuple, g[tag + "_bounds"])), tuple(int(r) for r in g[tag + "_res"]))
        fc = rp.flat_coords(grid)
        sel = rng.choice(len(fc), 250, replace=False)
        out, simp = linear_interpolate(g[tag + "_points"], g[tag + "_values"], fc[sel])
        assert np.array_equal(simp, g[tag + "_simplex"][sel])
        assert np.abs(out - g[tag + "_uvw"].reshape(3, -1).T[sel]).max() <= 1e-12
    cb = np.load(os.path.join(golden_dir, "case_b_boundary.npz"))
    grid, _ = rp.create_grid(((0, 12), (0, 12), (0, 12)), 12)
    fc = rp.flat_coords(grid)
    sel = rng.choice(len(fc), 250, replace=False)
    ctr = cb["points"].mean(0)
    q = fc[sel] + 2.0 ** -36 * (ctr - fc[sel])  # the kernel's nudge off the lattice planes
    out, _ = linear_interpolate(cb["points"], cb["values"], q)
    assert np.abs(out - g["c_uvw"].reshape(3, -1).T[sel]).max() <= 1e-8


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_linear_programme_restatement_vs_scipy_random_clouds(seed):
    """The SciPy-free restatement against scipy.spatial.Delaunay on fresh clouds: queries inside, just
    inside / outside the hull (on segments through hull vertices) and far outside."""
    from scipy.spatial import Delaunay
    from oracle.delaunay_lp import containing_simplex
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0, 5, size=(150, 3)) * np.array([1.0, 2.0, 0.5])  # anisotropic box
    tri = Delaunay(pts)
    hull_v = np.unique(tri.convex_hull)
    ctr = pts.mean(0)
    t = rng.uniform(0.9, 1.1, size=60)[:, None]
    near = ctr + t * (pts[rng.choice(hull_v, 60)] - ctr)           # around the hull along rays from the centre
    q = np.concatenate([rng.uniform(-1, 6, size=(80, 3)) * np.array([1.0, 2.0, 0.5]), near,
                        rng.uniform(20, 30, size=(5, 3))])
    s = tri.find_simplex(q)
    for qi, si in zip(q, s):
        ids, lam = containing_simplex(pts, qi)
        if si < 0:
            assert ids is None
        else:
            assert ids is not None and np.array_equal(ids, np.sort(tri.simplices[si]))
            assert abs(lam.sum() - 1.0) <= 1e-12 and lam.min() >= -1e-9
            assert np.abs(lam @ pts[ids] - qi).max() <= 1e-9


@pytest.mark.parametrize("kind", ["random", "lattice_faces", "sphere_shell"])
def test_hull_candidate_rule_keeps_every_hull_vertex(kind):
    """The closed-octant rule that prunes the kernel's hull-candidate list never drops a vertex of the convex
    hull (scipy.spatial.ConvexHull), also with co-planar lattice points on the faces, and prunes hard."""
    from scipy.spatial import ConvexHull
    from oracle.delaunay_lp import extreme_point_candidates
    rng = np.random.default_rng(5)
    if kind == "random":
        pts = rng.uniform(0, 10, size=(1500, 3))
    elif kind == "lattice_faces":
        g = np.arange(0, 9.0)
        face = np.array([(x, y, 0.0) for x in g for y in g] + [(x, y, 8.0) for x in g for y in g]
                        + [(0.0, y, z) for y in g for z in g] + [(8.0, y, z) for y in g for z in g])
        pts = np.concatenate([np.unique(face, axis=0), rng.uniform(0.5, 7.5, size=(600, 3))])
    else:
        v = rng.normal(size=(800, 3))
        pts = np.concatenate([5.0 * v / np.linalg.norm(v, axis=1, keepdims=True), rng.uniform(-2, 2, size=(400, 3))])
    keep = extreme_point_candidates(pts)
    hull = np.unique(ConvexHull(pts).vertices)
    assert keep[hull].all()
    if kind == "random":
        assert keep.sum() <= 0.2 * len(pts)
    if kind == "lattice_faces":  # the interior of the flat faces goes, only their outline can stay
        assert keep.sum() <= 8 * 9 + 40


def test_exact_division_by_spacing_cpu_restatement():
    """oracle/div_exact.c restates the stencil kernels' division by a grid spacing (reciprocal + two exact-residual FMA
    corrections, csrc/bulk_pipe.cuh) on the CPU: every quotient must be the IEEE one the reference computes
    (physics.py:26-53, np.gradient).  The GPU twin of this test is ptv_selftest_division."""
    import ctypes as C
    import subprocess
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    lib_path = os.path.join(here, "_build", "libdiv_exact.so")
    if not os.path.exists(lib_path):
        subprocess.run(["make", "-s", "-C", here], check=True)
    lib = C.CDLL(lib_path)
    lib.div_exact_mismatches.restype = C.c_int64
    lib.div_exact_mismatches.argtypes = [C.c_double, C.c_int64, C.c_uint64]
    for h in (2.00625, 4.0125, 1.25, 0.75, 3.0, 0.1, 1e-3, 7.0, -2.00625, 1.9999999999999998, 1.0000000000000002,
              1.0 / 3.0, 123456.789):
        assert lib.div_exact_mismatches(h, 2_000_000, 12345) == 0, h
