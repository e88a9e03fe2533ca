"""TEST INFRASTRUCTURE ONLY (CPU oracle) -- never imported by the product path.

Independent restatement of the simplex search behind ``griddata(method='linear')``
(reference: interpolator.py:197 -> scipy LinearNDInterpolator -> Qhull Delaunay + find_simplex;
SciPy is un-vendored and unpinned by the reference, 1.18.1 in this image).

For points in general position the Delaunay triangulation is unique, and the tetrahedron that
contains a query q is the optimum of a 4-variable linear programme: among all spheres with no data
point strictly inside, the one that holds q deepest (largest r^2 - |q-c|^2) is the circumsphere of
that tetrahedron.  ``containing_simplex`` solves the programme by brute-force dual-simplex pivoting
over ALL points (no spatial index, no SciPy): start from a huge tetrahedron of four virtual points
around q, repeatedly bring in the point deepest inside the current circumsphere and drop the vertex
chosen by the ratio test that keeps q inside.  ``tests/test_oracle_golden.py`` checks it against
``scipy.spatial.Delaunay.find_simplex`` and against golden vectors of the unmodified reference.

The CUDA kernel (csrc/delaunay_linear.cu) solves the same programme restricted to hash cells.  It picks
the entering point by violation per squared distance from q instead of the deepest one; the choice only
changes the path, not the optimum (every pivot raises the objective, the optimum is unique for points in
general position), which is why this restatement can keep the textbook rule.
"""
from __future__ import annotations

import numpy as np

VIRTUAL_DIRS = np.array([[1.0, 1.1, 0.9], [1.05, -1.0, -0.95], [-1.0, 0.93, -1.07], [-0.97, -1.02, 1.01]])


def _geom(v):
    e = v[1:] - v[0]
    det = np.dot(e[0], np.cross(e[1], e[2]))
    rows = np.stack([np.cross(e[1], e[2]), np.cross(e[2], e[0]), np.cross(e[0], e[1])]) / det
    h = 0.5 * np.einsum("ij,ij->i", e, e)
    c = h @ rows  # circumcentre relative to v[0]
    return rows, c


def _bary(v, rows, x):
    b = rows @ (x - v[0])
    return np.concatenate([[1.0 - b.sum()], b])


def containing_simplex(points, q, big=1e4, max_pivots=500):
    """Vertex indices (sorted) of the Delaunay tetrahedron of ``points`` containing ``q`` and the
    barycentric coordinates of q in that order; (None, None) if q is outside the convex hull."""
    points = np.asarray(points, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    extent = float(np.max(points.max(0) - points.min(0))) + float(np.max(np.abs(q - points.mean(0))))
    v = q + big * extent * VIRTUAL_DIRS
    ids = [-1, -2, -3, -4]
    for _ in range(max_pivots):
        rows, c = _geom(v)
        d = points - v[0]
        viol = 2.0 * (d @ c) - np.einsum("ij,ij->i", d, d)
        for i in ids:
            if i >= 0:
                viol[i] = -np.inf
        j = int(np.argmax(viol))
        if not viol[j] > 1e-12 * max(np.dot(c, c), 1e-300):
            break
        lam = np.maximum(_bary(v, rows, q), 0.0)
        mu = _bary(v, rows, points[j])
        ratio = np.where(mu > 1e-14, lam / np.where(mu > 1e-14, mu, 1.0), np.inf)
        out = int(np.argmin(ratio))
        v[out] = points[j]
        ids[out] = j
        real = [i for i in range(4) if ids[i] >= 0]
        if ids[0] < 0 and real:  # keep a real vertex as the reference point (small magnitudes)
            r = real[0]
            v[[0, r]] = v[[r, 0]]
            ids[0], ids[r] = ids[r], ids[0]
    else:
        raise RuntimeError("pivot limit")
    if min(ids) < 0:
        return None, None
    rows, _ = _geom(v)
    lam = _bary(v, rows, q)
    order = np.argsort(ids)
    return np.asarray(ids)[order], lam[order]


def linear_interpolate(points, values, queries, fill_value=0.0):
    """griddata(points, values, queries, method='linear', fill_value=...) by the programme above."""
    values = np.asarray(values, dtype=np.float64)
    out = np.full((len(queries), values.shape[1]), fill_value, dtype=np.float64)
    simp = np.full((len(queries), 4), -1, dtype=np.int64)
    for n, q in enumerate(np.asarray(queries, dtype=np.float64)):
        ids, lam = containing_simplex(points, q)
        if ids is not None:
            simp[n] = ids
            out[n] = lam @ values[ids]
    return out, simp


def extreme_point_candidates(points):
    """Restatement of the rule behind the kernel's hull-candidate list (csrc/delaunay_linear.cu,
    hull_refine_kernel): a point can be an extreme point of the cloud only if some CLOSED octant around it
    holds no other point.  (If every closed octant {s_i (x_i' - x_i) >= 0} held another point, any direction n
    would have n.(p' - p) >= 0 for the p' in the octant of sign(n), so p could not be strictly separated from
    the rest, i.e. it lies in their hull.)  O(N^2), for tests: returns the boolean candidate mask."""
    p = np.asarray(points, dtype=np.float64)
    d = p[None, :, :] - p[:, None, :]                      # d[i, j] = p_j - p_i
    distinct = np.any(d != 0.0, axis=2)
    keep = np.zeros(len(p), dtype=bool)
    for sx in (-1.0, 1.0):
        for sy in (-1.0, 1.0):
            for sz in (-1.0, 1.0):
                inside = (sx * d[..., 0] >= 0) & (sy * d[..., 1] >= 0) & (sz * d[..., 2] >= 0) & distinct
                keep |= ~inside.any(axis=1)
    return keep
