"""NumPy/SciPy restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference ``file:line`` it follows (paths are relative to the
upstream tree, ``tombultreys/ptv_interpolation``).  The neighbour search and the local RBF
solve live in un-vendored, unpinned SciPy (``requirements.txt:1-5``); the reference is
therefore "these scripts on the SciPy installed in this image (1.18.1)" and the port calls
the same SciPy entry points the reference calls, plus an independent brute-force kNN
(``knn_bruteforce``) that restates the published definition (exact Euclidean k nearest,
ascending distance).

Tie rule.  cKDTree returns equal-distance neighbours in tree-traversal order (it changes
with ``leafsize``), so there is no rule to follow.  The canonical order used on both sides
is ``(d2 ascending, particle index ascending)`` with ``d2 = (dx*dx + dy*dy) + dz*dz`` in
float64; ``knn_canonical`` over-queries cKDTree until the k-th distance is strictly inside
the queried set, recomputes d2 from coordinates and re-sorts.

Parity: pinned against the live reference by ``oracle/gen_golden.py`` ->
``tests/golden/*.npz`` (checked in ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "create_grid", "flat_coords", "knn_bruteforce", "knn_canonical", "idw_from_knn",
    "sibson_from_knn", "interpolate_field", "sample_mask_on_grid", "nearest_axis_index",
    "extract_boundary_particles", "compute_consistent_divergence", "flux_xy", "flux_xz",
    "flux_yz", "mid_plane_x_flux", "mean_abs_div", "apply_mask_zero", "outlier_keep_mask", "compute_strain_rate", "compute_vorticity",
    "build_laplacian_matrix", "apply_consistent_correction", "clean_divergence_projection",
    "delaunay_simplex_rows",
]


# --------------------------------------------------------------------------- grid
def create_grid(bounds, resolution):
    """interpolator.py:41-60 -- axes ``linspace(min, max-1, n)``, meshgrid over (z,y,x)."""
    (xmin, xmax), (ymin, ymax), (zmin, zmax) = bounds
    if isinstance(resolution, int):
        nx = ny = nz = resolution
    else:
        nx, ny, nz = resolution
    x = np.linspace(xmin, xmax - 1, nx)
    y = np.linspace(ymin, ymax - 1, ny)
    z = np.linspace(zmin, zmax - 1, nz)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    return (X, Y, Z), (x, y, z)


def flat_coords(grid_tuple):
    """interpolator.py:93/135/170 -- (Nvox,3) float64, C order: z slowest, x fastest."""
    X, Y, Z = grid_tuple
    return np.stack([np.ravel(X), np.ravel(Y), np.ravel(Z)], axis=-1)


# --------------------------------------------------------------------------- kNN
def _d2(q, p):
    """float64 squared distance with the summation order NumPy's ``.sum(-1)`` uses for
    three terms, ``(dx*dx + dy*dy) + dz*dz`` -- reproduces cKDTree's distances bit for
    bit after sqrt (measured)."""
    d = q - p
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def knn_bruteforce(points, queries, k, chunk=2048):
    """Definition of the search at interpolator.py:97,139 (``KDTree.query(q, k)``, p=2,
    eps=0): the k particles of smallest Euclidean distance, ascending; ties canonicalised
    by particle index.  O(Nq*Np); for small cases only."""
    points = np.ascontiguousarray(points, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    nq, n = len(queries), len(points)
    if k > n:
        raise IndexError(f"k={k} exceeds the number of particles {n}")
    idx_out = np.empty((nq, k), dtype=np.int64)
    d2_out = np.empty((nq, k), dtype=np.float64)
    ar = np.arange(n, dtype=np.int64)
    for s in range(0, nq, chunk):
        q = queries[s:s + chunk]
        d2 = _d2(q[:, None, :], points[None, :, :])
        # lexsort: last key is primary
        order = np.lexsort((np.broadcast_to(ar, d2.shape), d2), axis=1)[:, :k]
        idx_out[s:s + chunk] = order
        d2_out[s:s + chunk] = np.take_along_axis(d2, order, axis=1)
    return np.sqrt(d2_out), idx_out, d2_out


def knn_canonical(points, queries, k, workers=1, tree=None):
    """cKDTree search exactly as the reference calls it (interpolator.py:90,97,132,139;
    SciPy defaults leafsize=10, balanced_tree, compact_nodes) followed by the canonical
    re-ordering ``(d2, index)``.  Returns (dist, idx, d2), each (Nq,k)."""
    from scipy.spatial import KDTree
    points = np.ascontiguousarray(points, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    n = len(points)
    if k > n:
        # interpolator.py:150 `values[indices, i]` with index == Np -> IndexError
        raise IndexError(f"index {n} is out of bounds for axis 0 with size {n}")
    if tree is None:
        tree = KDTree(points)
    extra = min(n - k, 4)
    while True:
        kk = k + extra
        dist, idx = tree.query(queries, k=kk, workers=workers)
        if kk == 1:
            dist, idx = dist[:, None], idx[:, None]
        if kk == n or np.all(dist[:, kk - 1] > dist[:, k - 1]):
            break
        extra = min(n - k, max(2 * extra, 8))
    d2 = _d2(queries[:, None, :], points[idx])
    order = np.lexsort((idx, d2), axis=1)[:, :k]
    idx = np.take_along_axis(idx, order, axis=1)
    d2 = np.take_along_axis(d2, order, axis=1)
    return np.sqrt(d2), idx.astype(np.int64), d2


# --------------------------------------------------------------------------- weights
def idw_from_knn(distances, indices, values, idw_power=2.0):
    """interpolator.py:141-153 -- w = 1/(d**p + 1e-10); normalise; weighted sum per comp."""
    epsilon = 1e-10
    weights = 1.0 / (distances ** idw_power + epsilon)
    weights_sum = weights.sum(axis=1, keepdims=True)
    weights_normalized = weights / weights_sum
    out = np.zeros((len(distances), 3))
    for i in range(3):
        out[:, i] = (weights_normalized * values[indices, i]).sum(axis=1)
    return out


def sibson_from_knn(distances, indices, values):
    """interpolator.py:102-122 -- inverse-distance weights damped by exp(-d/std(d))."""
    epsilon = 1e-10
    inv_dist = 1.0 / (distances + epsilon)
    weights = inv_dist / inv_dist.sum(axis=1, keepdims=True)
    dist_std = distances.std(axis=1, keepdims=True)
    smoothing_factor = np.exp(-distances / (dist_std + epsilon))
    weights = weights * smoothing_factor
    weights = weights / weights.sum(axis=1, keepdims=True)
    out = np.zeros((len(distances), 3))
    for i in range(3):
        out[:, i] = (weights * values[indices, i]).sum(axis=1)
    return out


def _rbf_worker(interp, chunk):
    """interpolator.py:62-63"""
    return interp(chunk)


def interpolate_field(points, values, grid_tuple, method="idw", rbf_neighbors=20,
                      rbf_kernel="thin_plate_spline", smoothing=0.0, idw_power=2.0,
                      idw_neighbors=50, sibson_neighbors=30, workers=1, chunk_voxels=1 << 18,
                      canonical=True, return_knn=False, n_jobs=1):
    """interpolator.py:65-203 for methods idw / sibson / rbf / nearest, driven in voxel
    chunks so large grids fit in RAM (per-voxel results do not depend on the chunking:
    the tree always holds every particle).  ``points``/``values`` are the (Np,3) float64
    arrays of interpolator.py:78-79.  ``canonical=False`` uses cKDTree's own tie order
    (what the reference literally does; used for CPU timing)."""
    from scipy.spatial import KDTree
    points = np.ascontiguousarray(points, dtype=np.float64)
    values = np.ascontiguousarray(values, dtype=np.float64)
    X = grid_tuple[0]
    fc = flat_coords(grid_tuple)
    nq = len(fc)
    out = np.zeros((nq, 3))
    if method in ("idw", "sibson"):
        k = idw_neighbors if method == "idw" else sibson_neighbors
        if k > len(points):
            raise IndexError(f"index {len(points)} is out of bounds for axis 0 with size {len(points)}")
        tree = KDTree(points)
        keep_d, keep_i = [], []
        for s in range(0, nq, chunk_voxels):
            q = fc[s:s + chunk_voxels]
            if canonical:
                dist, idx, _ = knn_canonical(points, q, k, workers=workers, tree=tree)
            else:
                dist, idx = tree.query(q, k=k, workers=workers)
                if k == 1:
                    dist, idx = dist[:, None], idx[:, None]
            if method == "idw":
                out[s:s + chunk_voxels] = idw_from_knn(dist, idx, values, idw_power)
            else:
                out[s:s + chunk_voxels] = sibson_from_knn(dist, idx, values)
            if return_knn:
                keep_d.append(dist)
                keep_i.append(idx)
    elif method == "rbf":
        # interpolator.py:162-167,190 -> scipy.interpolate.RBFInterpolator (local mode)
        from scipy.interpolate import RBFInterpolator
        interp = RBFInterpolator(points, values, neighbors=rbf_neighbors, kernel=rbf_kernel,
                                 smoothing=smoothing)
        if n_jobs > 1:
            # interpolator.py:173-182 (the test_parallel.py mode): the voxels are split into n_jobs
            # contiguous chunks, the interpolator is pickled to every worker, results are stacked
            from concurrent.futures import ProcessPoolExecutor
            chunks = np.array_split(fc, n_jobs)
            with ProcessPoolExecutor(max_workers=n_jobs) as executor:
                results = list(executor.map(_rbf_worker, [interp] * n_jobs, chunks))
            out[:] = np.vstack(results)
        else:
            for s in range(0, nq, 10000):
                out[s:s + 10000] = interp(fc[s:s + 10000])
    elif method == "nearest":
        # interpolator.py:197 griddata(method='nearest') == cKDTree k=1 lookup
        dist, idx, _ = knn_canonical(points, fc, 1, workers=workers)
        out[:] = values[idx[:, 0]]
    elif method == "linear":
        # interpolator.py:197 griddata(points, values, grid, method='linear', fill_value=0.0)
        # == scipy LinearNDInterpolator (Qhull Delaunay, find_simplex, barycentric weights).  An
        # independent restatement without SciPy is oracle/delaunay_lp.py.
        from scipy.interpolate import LinearNDInterpolator
        interp = LinearNDInterpolator(points, values, fill_value=0.0)
        for s in range(0, nq, chunk_voxels):
            out[s:s + chunk_voxels] = interp(fc[s:s + chunk_voxels])
    else:
        raise NotImplementedError(method)
    interp3 = out.reshape(X.shape + (3,))
    U, V, W = interp3[..., 0], interp3[..., 1], interp3[..., 2]
    if return_knn and method in ("idw", "sibson"):
        return U, V, W, np.concatenate(keep_d), np.concatenate(keep_i)
    return U, V, W


def delaunay_simplex_rows(points, queries):
    """Vertex rows (ascending) of the tetrahedron ``scipy.spatial.Delaunay(points).find_simplex`` returns
    for each query and the barycentric weights in that order; rows are -1 outside the convex hull.
    This is the simplex LinearNDInterpolator evaluates in (interpolator.py:197)."""
    from scipy.spatial import Delaunay
    points = np.ascontiguousarray(points, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    tri = Delaunay(points)
    s = tri.find_simplex(queries)
    ok = s >= 0
    simp = tri.simplices[np.maximum(s, 0)]
    T = tri.transform[np.maximum(s, 0)]
    b3 = np.einsum("nij,nj->ni", T[:, :3, :], queries - T[:, 3, :])
    b = np.concatenate([b3, 1.0 - b3.sum(1, keepdims=True)], axis=1)
    order = np.argsort(simp, axis=1)
    rows = np.take_along_axis(simp, order, 1).astype(np.int64)
    b = np.take_along_axis(b, order, 1)
    rows[~ok] = -1
    b[~ok] = np.nan
    return rows, b


def apply_mask_zero(U, V, W, mask):
    """main.py:195-207 -- NaN -> 0 then hard zero in solid voxels (mask False)."""
    outs = []
    for a in (U, V, W):
        a = np.nan_to_num(np.array(a, copy=True))
        a[~mask] = 0
        outs.append(a)
    return tuple(outs)


# --------------------------------------------------------------------------- projection cleaning (N2)
def build_laplacian_matrix(mask, dx, dy, dz):
    """physics.py:55-108 -- masked 7-point Laplacian over fluid voxels as a CSR matrix + index map."""
    from scipy import sparse
    nz, ny, nx = mask.shape
    n_fluid = np.sum(mask)
    idx_map = np.full(mask.shape, -1, dtype=np.int32)
    idx_map[mask] = np.arange(n_fluid)
    rows, cols, data = [], [], []
    I, J, K = np.where(mask)
    curr = idx_map[I, J, K]
    for axis, h2_inv in [(2, 1.0 / (dx**2)), (1, 1.0 / (dy**2)), (0, 1.0 / (dz**2))]:
        for offset in [-1, 1]:
            In, Jn, Kn = I, J, K
            if axis == 2:
                Kn = K + offset
            elif axis == 1:
                Jn = J + offset
            else:
                In = I + offset
            valid_b = (In >= 0) & (In < nz) & (Jn >= 0) & (Jn < ny) & (Kn >= 0) & (Kn < nx)
            neigh = np.full_like(curr, -1)
            neigh[valid_b] = idx_map[In[valid_b], Jn[valid_b], Kn[valid_b]]
            connected = neigh != -1
            rows.append(curr[connected]); cols.append(neigh[connected])
            data.append(np.full(np.sum(connected), h2_inv))
            rows.append(curr[connected]); cols.append(curr[connected])
            data.append(np.full(np.sum(connected), -h2_inv))
    A = sparse.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(n_fluid, n_fluid))
    return A.tocsr(), idx_map


def apply_consistent_correction(u, v, w, phi, mask, dx, dy, dz):
    """physics.py:110-147 -- subtract the cell average of the staggered face gradients of phi."""
    phi_grid = np.zeros_like(u)
    phi_grid[mask] = phi

    def cell_grad(p, axis, m, h):
        p = np.moveaxis(p, axis, -1)
        m = np.moveaxis(m, axis, -1)
        g_next = np.zeros_like(p)
        g_next[..., :-1] = np.where(m[..., 1:] & m[..., :-1], (p[..., 1:] - p[..., :-1]) / h, 0.0)
        g_prev = np.zeros_like(p)
        g_prev[..., 1:] = g_next[..., :-1]
        return np.moveaxis((g_next + g_prev) / 2.0, -1, axis)

    u_new = u - cell_grad(phi_grid, 2, mask, dx)
    v_new = v - cell_grad(phi_grid, 1, mask, dy)
    w_new = w - cell_grad(phi_grid, 0, mask, dz)
    u_new[~mask] = 0
    v_new[~mask] = 0
    w_new[~mask] = 0
    return u_new, v_new, w_new


def clean_divergence_projection(u, v, w, mask, dx, dy, dz, iterations=3, return_phi=False):
    """physics.py:149-209 without the prints: divergence -> b - mean(b) -> lsqr(damp=1e-8,
    atol=btol=1e-10, iter_lim=3000) -> correction, `iterations` times."""
    from scipy.sparse.linalg import lsqr
    u_c, v_c, w_c = u.copy(), v.copy(), w.copy()
    phis = []
    for i in range(iterations):
        div = compute_consistent_divergence(u_c, v_c, w_c, mask, dx, dy, dz)
        if i == 0:
            A, _ = build_laplacian_matrix(mask, dx, dy, dz)
        b = div[mask]
        b = b - np.mean(b)
        res = lsqr(A, b, damp=1e-8, atol=1e-10, btol=1e-10, iter_lim=3000, show=False)
        phi = res[0]
        phis.append(res)
        if np.isnan(phi).any():
            break
        u_c, v_c, w_c = apply_consistent_correction(u_c, v_c, w_c, phi, mask, dx, dy, dz)
    if return_phi:
        return u_c, v_c, w_c, phis
    return u_c, v_c, w_c


# --------------------------------------------------------------------------- gradient stencils (N3)
def compute_strain_rate(u, v, w, dx, dy, dz, mask=None):
    """velocity_analysis.py:10-63."""
    du_dz, du_dy, du_dx = np.gradient(u, dz, dy, dx)
    dv_dz, dv_dy, dv_dx = np.gradient(v, dz, dy, dx)
    dw_dz, dw_dy, dw_dx = np.gradient(w, dz, dy, dx)
    exx, eyy, ezz = 2 * du_dx, 2 * dv_dy, 2 * dw_dz
    exy, exz, eyz = du_dy + dv_dx, du_dz + dw_dx, dv_dz + dw_dy
    s = np.sqrt(0.5 * (exx**2 + eyy**2 + ezz**2) + exy**2 + exz**2 + eyz**2)
    if mask is not None:
        s[~mask] = 0.0
    return s


def compute_vorticity(u, v, w, dx, dy, dz, mask=None):
    """velocity_analysis.py:94-120."""
    du_dz, du_dy, du_dx = np.gradient(u, dz, dy, dx)
    dv_dz, dv_dy, dv_dx = np.gradient(v, dz, dy, dx)
    dw_dz, dw_dy, dw_dx = np.gradient(w, dz, dy, dx)
    o = np.sqrt((dw_dy - dv_dz)**2 + (du_dz - dw_dx)**2 + (dv_dx - du_dy)**2)
    if mask is not None:
        o[~mask] = 0.0
    return o


# --------------------------------------------------------------------------- outlier filter (N1)
def outlier_keep_mask(points, values, k=25, threshold=3.0, workers=1):
    """filtering.py:5-58 restated: returns (keep_mask, kth_dist).  The self-query's first column is
    dropped in the canonical (d2, index) order."""
    points = np.ascontiguousarray(points, dtype=np.float64)
    u, v, w = values[:, 0], values[:, 1], values[:, 2]
    speed = np.sqrt(u**2 + v**2 + w**2)
    dist, idx, _ = knn_canonical(points, points, k + 1, workers=workers)
    neighbor_indices = idx[:, 1:]
    neighbor_distances = dist[:, 1:]
    neighbor_speeds = speed[neighbor_indices]
    local_medians = np.median(neighbor_speeds, axis=1)
    local_mads = np.median(np.abs(neighbor_speeds - local_medians[:, np.newaxis]), axis=1)
    z_scores = (np.abs(speed - local_medians)) / (local_mads + 1e-6)
    return z_scores <= threshold, neighbor_distances[:, -1]


# --------------------------------------------------------------------------- mask
def nearest_axis_index(src_coords, q):
    """Per-axis restatement of RegularGridInterpolator(method='nearest',
    bounds_error=False) as used at interpolator.py:226-236: interval search
    (scipy/interpolate/_rgi.py:632-633 -> find_indices), normalised offset
    ``(q-g[i])/(g[i+1]-g[i])``, ``<= 0.5 -> i else i+1`` (_rgi.py:551-554); -1 marks
    out of bounds (_rgi.py:635-642: q < g[0] or q > g[-1])."""
    g = np.asarray(src_coords, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    n = len(g)
    if n == 1:
        idx = np.zeros(q.shape, dtype=np.int64)
    else:
        i = np.clip(np.searchsorted(g, q, side="right") - 1, 0, n - 2)
        yi = (q - g[i]) / (g[i + 1] - g[i])
        idx = np.where(yi <= 0.5, i, i + 1).astype(np.int64)
    oob = (q < g[0]) | (q > g[-1])
    idx[oob] = -1
    return idx


def sample_mask_on_grid(mask_raw, grid_tuple, bounds_raw, use_scipy=True):
    """interpolator.py:205-238.  ``use_scipy=True`` calls the same SciPy interpolator the
    reference calls; ``False`` uses the separable restatement above (what the CUDA path
    implements).  Both must agree bit for bit (tests/test_oracle_golden.py)."""
    nz, ny, nx = mask_raw.shape
    (xmin, xmax), (ymin, ymax), (zmin, zmax) = bounds_raw
    X, Y, Z = grid_tuple
    z_coords = np.linspace(zmin, zmax - 1, nz) if nz > 1 else np.array([zmin])
    y_coords = np.linspace(ymin, ymax - 1, ny) if ny > 1 else np.array([ymin])
    x_coords = np.linspace(xmin, xmax - 1, nx) if nx > 1 else np.array([xmin])
    if use_scipy:
        from scipy.interpolate import RegularGridInterpolator
        interp = RegularGridInterpolator((z_coords, y_coords, x_coords), mask_raw.astype(float),
                                         method="nearest", bounds_error=False, fill_value=0)
        pts = np.stack([np.ravel(Z), np.ravel(Y), np.ravel(X)], axis=-1)
        return interp(pts).reshape(X.shape) > 0.5
    iz = nearest_axis_index(z_coords, np.ravel(Z))
    iy = nearest_axis_index(y_coords, np.ravel(Y))
    ix = nearest_axis_index(x_coords, np.ravel(X))
    ok = (iz >= 0) & (iy >= 0) & (ix >= 0)
    out = np.zeros(iz.shape, dtype=bool)
    out[ok] = mask_raw[iz[ok], iy[ok], ix[ok]]
    return out.reshape(X.shape)


def extract_boundary_particles(mask, bounds, sampling_step=1, thickness=1):
    """interpolator.py:240-284.  ``thickness`` iterations of 6-connected binary dilation
    (scipy.ndimage, border_value=0) == solid voxels within Manhattan distance
    ``thickness`` of a fluid voxel; C-order (z,y,x) enumeration; every ``sampling_step``-th;
    index -> coordinate ``min + i*(max-1-min)/(n-1)``."""
    if mask is None:
        return np.array([]), np.array([]), np.array([])
    nz, ny, nx = mask.shape
    (xmin, xmax), (ymin, ymax), (zmin, zmax) = bounds
    cur = np.asarray(mask, dtype=bool)
    steps = thickness
    if thickness < 1:
        # scipy.ndimage.binary_dilation(iterations < 1) repeats until nothing changes: with 6-connectivity and
        # border_value = 0 that fills the whole box as soon as one voxel is fluid
        steps = (nz + ny + nx) if cur.any() else 0
    for _ in range(steps):
        nxt = cur.copy()
        nxt[1:, :, :] |= cur[:-1, :, :]
        nxt[:-1, :, :] |= cur[1:, :, :]
        nxt[:, 1:, :] |= cur[:, :-1, :]
        nxt[:, :-1, :] |= cur[:, 1:, :]
        nxt[:, :, 1:] |= cur[:, :, :-1]
        nxt[:, :, :-1] |= cur[:, :, 1:]
        cur = nxt
    boundary = cur & (~np.asarray(mask, dtype=bool))
    Z_idx, Y_idx, X_idx = np.where(boundary)
    if len(X_idx) == 0:
        return np.array([]), np.array([]), np.array([])
    if sampling_step > 1:
        Z_idx, Y_idx, X_idx = Z_idx[::sampling_step], Y_idx[::sampling_step], X_idx[::sampling_step]
    z_phys = zmin + Z_idx * (zmax - 1 - zmin) / (nz - 1) if nz > 1 else np.full_like(Z_idx, zmin)
    y_phys = ymin + Y_idx * (ymax - 1 - ymin) / (ny - 1) if ny > 1 else np.full_like(Y_idx, ymin)
    x_phys = xmin + X_idx * (xmax - 1 - xmin) / (nx - 1) if nx > 1 else np.full_like(X_idx, xmin)
    return x_phys, y_phys, z_phys


# --------------------------------------------------------------------------- stencils
def _face_diff(f, m, axis):
    """Closed form of physics.py:26-47 along one axis:
    F+[i] = m[i+1] ? (f[i]+f[i+1])/2 : 0 (i<n-1), F+[n-1]=f[n-1]; F-[i]=F+[i-1], F-[0]=f[0]."""
    f = np.moveaxis(np.asarray(f, dtype=np.float64), axis, -1)
    m = np.moveaxis(np.asarray(m, dtype=bool), axis, -1)
    fp = np.empty_like(f)
    fp[..., :-1] = np.where(m[..., 1:], (f[..., :-1] + f[..., 1:]) / 2.0, 0.0)
    fp[..., -1] = f[..., -1]
    fm = np.empty_like(f)
    fm[..., 1:] = fp[..., :-1]
    fm[..., 0] = f[..., 0]
    return np.moveaxis(fp - fm, -1, axis)


def compute_consistent_divergence(u, v, w, mask, dx, dy, dz):
    """physics.py:6-53 -- (F+ - F-)/h summed over x (u, axis 2), y (v, axis 1), z (w, axis 0)."""
    return _face_diff(u, mask, 2) / dx + _face_diff(v, mask, 1) / dy + _face_diff(w, mask, 0) / dz


def flux_xy(w_field, dx, dy):
    """plot_flux.py:6-8."""
    return np.sum(w_field, axis=(1, 2)) * dx * dy


def flux_xz(v_field, dx, dz):
    """plot_flux.py:10-12."""
    return np.sum(v_field, axis=(0, 2)) * dx * dz


def flux_yz(u_field, dy, dz):
    """plot_flux.py:14-16."""
    return np.sum(u_field, axis=(0, 1)) * dy * dz


def mid_plane_x_flux(u_field, dy, dz):
    """physics.py:160-165."""
    nx = u_field.shape[2]
    return np.sum(u_field[:, :, nx // 2]) * dy * dz


def mean_abs_div(div, mask):
    """physics.py:174 / view_divergence.py:45."""
    return np.mean(np.abs(div[mask]))
