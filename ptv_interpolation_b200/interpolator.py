"""Drop-in mirror of the reference's ``interpolator.py`` public API, backed by the sm_100a
CUDA path (libptvb200.so).  Same names, argument meaning and error behaviour:

    load_ptv_data, load_mask, create_grid, interpolate_field, sample_mask_on_grid,
    extract_boundary_particles                       (reference interpolator.py:9,28,41,65,205,240)

Differences a caller can observe (all documented in DESIGN.md):
  * ``create_grid`` returns read-only zero-stride broadcast views for X, Y, Z (same values and
    shapes; 24 B/voxel of coordinates are never materialised);
  * ``interpolate_field`` returns float32 arrays by default (``out_dtype=np.float64`` restores the
    reference dtype), accepts an optional ``mask=`` to skip and zero solid voxels (what
    main.py:202-207 does afterwards), and writes 0 where the reference would produce NaN
    (main.py:195-199 replaces those by 0 anyway);
  * ``method='linear'`` (griddata / Qhull Delaunay, interpolator.py:197 -- the reference's default) runs
    on the CUDA path without building a triangulation (csrc/delaunay_linear.cu); ``'cubic'`` raises the
    ValueError griddata raises for 3-D data.
"""
from __future__ import annotations

import numpy as np

from .engine import default_engine

__all__ = ["load_ptv_data", "load_mask", "create_grid", "interpolate_field", "sample_mask_on_grid",
           "extract_boundary_particles"]

_GPU_METHODS = ("linear", "idw", "sibson", "nearest", "rbf")


def load_ptv_data(filepath):
    """interpolator.py:9-26 -- CSV with x,y,z,u,v,w (or vx,vy,vz); any failure -> IOError."""
    try:
        import pandas as pd
        df = pd.read_csv(filepath)
        df.rename(columns={"vx": "u", "vy": "v", "vz": "w"}, inplace=True)
        required_cols = {"x", "y", "z", "u", "v", "w"}
        if not required_cols.issubset(df.columns):
            raise ValueError(f"CSV must contain columns: {required_cols}")
        return df
    except Exception as e:
        raise IOError(f"Error reading {filepath}: {e}")


def load_mask(filepath):
    """interpolator.py:28-39 -- 3-D TIFF (or .npy when tifffile is unavailable) -> bool, True = fluid."""
    try:
        if str(filepath).endswith(".npy"):
            mask = np.load(filepath)
        else:
            import tifffile
            mask = tifffile.imread(filepath)
        return mask > 0
    except Exception as e:
        raise IOError(f"Error reading mask {filepath}: {e}")


def create_grid(bounds, resolution):
    """interpolator.py:41-60 -- axes ``linspace(min, max-1, n)``; X, Y, Z of shape (nz, ny, nx)."""
    (xmin, xmax), (ymin, ymax), (zmin, zmax) = bounds
    if isinstance(resolution, int):
        nx = ny = nz = resolution
    else:
        nx, ny, nz = resolution
    x = np.linspace(xmin, xmax - 1, nx)
    y = np.linspace(ymin, ymax - 1, ny)
    z = np.linspace(zmin, zmax - 1, nz)
    shape = (nz, ny, nx)
    X = np.broadcast_to(x[None, None, :], shape)
    Y = np.broadcast_to(y[None, :, None], shape)
    Z = np.broadcast_to(z[:, None, None], shape)
    return (X, Y, Z), (x, y, z)


def _grid_axes(grid_tuple):
    """Recover the three float64 axes of a rectilinear (nz,ny,nx) meshgrid; zero-stride views are
    recognised without touching memory, dense meshgrids are verified."""
    X, Y, Z = (np.asarray(a) for a in grid_tuple)
    if X.ndim != 3 or X.shape != Y.shape or X.shape != Z.shape:
        raise ValueError("grid_tuple must hold three (nz, ny, nx) arrays")
    x = np.ascontiguousarray(X[0, 0, :], dtype=np.float64)
    y = np.ascontiguousarray(Y[0, :, 0], dtype=np.float64)
    z = np.ascontiguousarray(Z[:, 0, 0], dtype=np.float64)

    def is_bcast(a, axis):
        return all(a.strides[i] == 0 or a.shape[i] == 1 for i in range(3) if i != axis)

    if not (is_bcast(X, 2) and is_bcast(Y, 1) and is_bcast(Z, 0)):
        ok = (np.array_equal(X, np.broadcast_to(X[0:1, 0:1, :], X.shape))
              and np.array_equal(Y, np.broadcast_to(Y[0:1, :, 0:1], Y.shape))
              and np.array_equal(Z, np.broadcast_to(Z[:, 0:1, 0:1], Z.shape)))
        if not ok:
            return None  # arbitrary query points: handled by the point-query kernel
    return x, y, z


def interpolate_field(df, grid_tuple, method="linear", rbf_neighbors=20, rbf_kernel="thin_plate_spline",
                      smoothing=0.0, n_jobs=1, idw_power=2.0, idw_neighbors=50, sibson_neighbors=30,
                      mask=None, out_dtype=np.float32, device=None, return_knn=False, shard_inputs=False):
    """interpolator.py:65-203.  ``n_jobs`` is accepted and ignored (one GPU does the work).
    Returns (U, V, W): three writable (nz, ny, nx) views of one host array, like the reference.
    ``shard_inputs=True`` (one process per GPU under torch.distributed, every rank calling with the SAME
    DataFrame and its own z-slab of the grid): each rank uploads 1/N of the particle table and the ranks
    all-gather it over NVLink instead of N full host->device copies."""
    import torch
    if method not in _GPU_METHODS:
        # interpolator.py:197 hands every other name to griddata, which knows 'cubic' only in 1-D / 2-D
        raise ValueError(f"Unknown interpolation method {method!r} for 3 dimensional data")
    if method == "rbf":
        from .engine import method_code
        method_code("rbf", rbf_kernel)  # RBFInterpolator's kernel / epsilon checks (ValueError)
    axes = _grid_axes(grid_tuple)
    eng = default_engine(device)
    dev = eng.device
    if axes is None:
        points = df[["x", "y", "z"]].values  # interpolator.py:78-79
        values = df[["u", "v", "w"]].values
        return _interpolate_scattered(eng, points, values, grid_tuple, method, rbf_neighbors, smoothing, idw_power,
                                      idw_neighbors, sibson_neighbors, mask, out_dtype, return_knn, rbf_kernel)
    x, y, z = axes
    from . import hostmem
    # particle table (interpolator.py:78-79): the six columns go to the device one by one through the pinned
    # staging chunks and are interleaved into (Np,3) rows THERE -- df[[...]].values would transpose 2 x 24 B
    # per particle on one host core first
    pts, vals = _tables_to_device(df, dev, shard_inputs)
    npart = pts.shape[0]
    k = {"idw": idw_neighbors, "sibson": sibson_neighbors, "nearest": 1, "rbf": rbf_neighbors, "linear": 4}[method]
    if method == "rbf":
        k = min(int(k), npart)  # scipy _rbfinterp.py:313 clamps silently
    eng.build(pts, vals)
    ax = [torch.from_numpy(np.array(a, dtype=np.float64)).to(dev) for a in (x, y, z)]
    m = mh = None
    if mask is not None:
        mh = np.ascontiguousarray(mask)
        if mh.dtype != np.bool_ and mh.dtype != np.uint8:
            mh = mh != 0
        if return_knn:
            m = hostmem.stage_to_device(mh, dev)
    tdt = torch.float32 if np.dtype(out_dtype) == np.float32 else torch.float64
    kw = dict(method=method, k=int(k), idw_power=float(idw_power), smoothing=float(smoothing), rbf_kernel=rbf_kernel)
    if return_knn:
        out, kd, ki = eng.interpolate(ax[0], ax[1], ax[2], mask=m, out_dtype=tdt, return_knn=True, **kw)
        host = out.cpu().numpy()
        return host[0], host[1], host[2], kd.cpu().numpy(), ki.cpu().numpy()
    # pinned result buffer from the pool (cudaHostAlloc of a 1024^3 result costs seconds), filled chunk by
    # chunk while later z-chunks are still being searched; it goes back to the pool when the caller has
    # dropped U, V and W
    shape = (3, len(z), len(y), len(x))
    host_t = hostmem.results.take(shape, tdt)
    dkey = (shape, tdt, dev.index)
    dev_out = _dev_results.pop(dkey, None)
    if dev_out is None:
        dev_out = torch.empty(shape, dtype=tdt, device=dev)
    # the mask enters z-chunk by z-chunk, each chunk staged while the previous one is being searched
    _, finished = eng.interpolate_to_host(ax[0], ax[1], ax[2], host_t, mask_host=mh, dev_out=dev_out, **kw)
    finished.synchronize()
    _dev_results.clear()  # keep at most one device result buffer alive between calls
    _dev_results[dkey] = dev_out
    host = hostmem.results.as_numpy(host_t)
    return host[0], host[1], host[2]


_dev_results = {}


def _tables_to_device(df, dev, shard=False):
    """(points, values): the two (Np,3) float64 device tensors of interpolator.py:78-79.  One rank: the six columns
    travel as ONE stream of staging chunks into a (6, Np) device buffer and are interleaved there."""
    import torch
    import torch.distributed as dist
    from . import hostmem
    world = dist.get_world_size() if (shard and dist.is_available() and dist.is_initialized()) else 1
    if world > 1:
        return _columns_to_device(df, ("x", "y", "z"), dev, True), _columns_to_device(df, ("u", "v", "w"), dev, True)
    n = len(df)
    cols = [np.ascontiguousarray(df[c].to_numpy(dtype=np.float64, copy=False)) for c in ("x", "y", "z", "u", "v", "w")]
    tmp = torch.empty((6, n), dtype=torch.float64, device=dev)
    hostmem.stage_many_to_device([(col, tmp[j]) for j, col in enumerate(cols)], dev)
    return tmp[:3].t().contiguous(), tmp[3:].t().contiguous()


def _columns_to_device(df, cols, dev, shard=False):
    """(Np, len(cols)) float64 device tensor of the DataFrame's columns, rows interleaved on the device.
    ``shard``: this rank uploads only its 1/N of the rows, the ranks all-gather the rest over NCCL."""
    import torch
    import torch.distributed as dist
    from . import hostmem
    n = len(df)
    out = torch.empty((n, len(cols)), dtype=torch.float64, device=dev)
    world = dist.get_world_size() if (shard and dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        for j, c in enumerate(cols):
            col = np.ascontiguousarray(df[c].to_numpy(dtype=np.float64, copy=False))
            out[:, j] = hostmem.stage_to_device(col, dev)
        return out
    rank = dist.get_rank()
    m = -(-n // world)  # rows per rank, the last shard is padded
    lo, hi = min(rank * m, n), min((rank + 1) * m, n)
    mine = torch.zeros((len(cols), m), dtype=torch.float64, device=dev)
    for j, c in enumerate(cols):
        col = np.ascontiguousarray(df[c].to_numpy(dtype=np.float64, copy=False)[lo:hi])
        if hi > lo:
            hostmem.stage_to_device(col, dev, out=mine[j, :hi - lo])
    gathered = torch.empty((world, len(cols), m), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gathered, mine)
    out.copy_(gathered.permute(0, 2, 1).reshape(world * m, len(cols))[:n])
    return out


def _interpolate_scattered(eng, points, values, grid_tuple, method, rbf_neighbors, smoothing, idw_power,
                           idw_neighbors, sibson_neighbors, mask, out_dtype, return_knn,
                           rbf_kernel="thin_plate_spline"):
    """interpolate_field for a grid_tuple that is not a rectilinear meshgrid: the (X, Y, Z) arrays are
    treated as arbitrary query points (what the reference does with every grid, interpolator.py:93)."""
    import torch
    from .engine import PTVEngine
    dev = eng.device
    X, Y, Z = (np.asarray(a, dtype=np.float64) for a in grid_tuple)
    shape = X.shape
    q = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=-1)
    sel = None
    if mask is not None:
        sel = np.flatnonzero(np.asarray(mask).ravel() != 0)
        q = q[sel]
    k = {"idw": idw_neighbors, "sibson": sibson_neighbors, "nearest": 1, "rbf": rbf_neighbors, "linear": 4}[method]
    if method == "rbf":
        k = min(int(k), len(points))
    tdt = torch.float32 if np.dtype(out_dtype) == np.float32 else torch.float64
    eng.build(torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)).to(dev),
              torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64)).to(dev))
    full = np.zeros((3,) + shape, dtype=out_dtype)
    kd = ki = None
    if len(q):
        qe = PTVEngine(dev)
        qt = torch.from_numpy(np.ascontiguousarray(q)).to(dev)
        qe.build(qt, qt)
        res = eng.interpolate_points(qe, method=method, k=int(k), idw_power=float(idw_power),
                                     smoothing=float(smoothing), out_dtype=tdt, return_knn=return_knn,
                                     rbf_kernel=rbf_kernel)
        out = res[0] if return_knn else res
        flat = full.reshape(3, -1)
        if sel is None:
            flat[:] = out.cpu().numpy()
        else:
            flat[:, sel] = out.cpu().numpy()
        if return_knn:
            kd, ki = res[1].cpu().numpy(), res[2].cpu().numpy()
        qe.close()
    if return_knn:
        return full[0], full[1], full[2], kd, ki
    return full[0], full[1], full[2]


def _nearest_axis_index(src_coords, q):
    """RegularGridInterpolator(method='nearest', bounds_error=False) along one axis
    (scipy _rgi.py:551-554, 632-642): interval i, offset (q-g[i])/(g[i+1]-g[i]), <= 0.5 -> i else
    i+1; -1 where q lies outside [g[0], g[-1]]."""
    g = np.asarray(src_coords, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    n = len(g)
    if n == 1:
        idx = np.zeros(q.shape, dtype=np.int64)
    else:
        i = np.clip(np.searchsorted(g, q, side="right") - 1, 0, n - 2)
        yi = (q - g[i]) / (g[i + 1] - g[i])
        idx = np.where(yi <= 0.5, i, i + 1).astype(np.int64)
    idx[(q < g[0]) | (q > g[-1])] = -1
    return idx.astype(np.int32)


def sample_mask_on_grid(mask_raw, grid_tuple, bounds_raw, device=None):
    """interpolator.py:205-238 -- nearest-neighbour resampling of ``mask_raw`` onto the grid."""
    import torch
    mask_raw = np.asarray(mask_raw)
    nz, ny, nx = mask_raw.shape
    (xmin, xmax), (ymin, ymax), (zmin, zmax) = bounds_raw
    axes = _grid_axes(grid_tuple)
    z_coords = np.linspace(zmin, zmax - 1, nz) if nz > 1 else np.array([zmin])
    y_coords = np.linspace(ymin, ymax - 1, ny) if ny > 1 else np.array([ymin])
    x_coords = np.linspace(xmin, xmax - 1, nx) if nx > 1 else np.array([xmin])
    eng = default_engine(device)
    dev = eng.device
    if axes is None:
        # not a rectilinear meshgrid: the reference samples whatever (X, Y, Z) holds point by point
        # (interpolator.py:233-236) -- nearest index per point and axis, then one device gather
        X, Y, Z = (np.asarray(a, dtype=np.float64) for a in grid_tuple)
        idx = [torch.from_numpy(_nearest_axis_index(c, q.ravel()).astype(np.int64)).to(dev)
               for c, q in ((x_coords, X), (y_coords, Y), (z_coords, Z))]
        raw = (mask_raw.astype(float) > 0.5) if mask_raw.dtype != np.bool_ else mask_raw
        raw_t = torch.from_numpy(np.ascontiguousarray(raw)).to(dev)
        inside = (idx[0] >= 0) & (idx[1] >= 0) & (idx[2] >= 0)
        got = raw_t[idx[2].clamp_min(0), idx[1].clamp_min(0), idx[0].clamp_min(0)] & inside
        return got.reshape(X.shape).cpu().numpy().astype(bool)
    x, y, z = axes
    ix = torch.from_numpy(_nearest_axis_index(x_coords, x)).to(dev)
    iy = torch.from_numpy(_nearest_axis_index(y_coords, y)).to(dev)
    iz = torch.from_numpy(_nearest_axis_index(z_coords, z)).to(dev)
    # mask_raw.astype(float) > 0.5 after nearest lookup == (value != 0) for bool / 0-1 masks;
    # general numeric masks are thresholded the same way the reference does (:238)
    raw = (mask_raw.astype(float) > 0.5) if mask_raw.dtype != np.bool_ else mask_raw
    raw_t = torch.from_numpy(np.ascontiguousarray(raw).view(np.uint8)).to(dev)
    out = eng.mask_gather(raw_t, ix, iy, iz)
    return out.cpu().numpy().astype(bool)


def extract_boundary_particles(mask, bounds, sampling_step=1, thickness=1, device=None):
    """interpolator.py:240-284 -- coordinates of solid voxels within ``thickness`` dilation steps
    of the fluid, every ``sampling_step``-th in C order, for zero-velocity wall particles."""
    import torch
    if mask is None:
        return np.array([]), np.array([]), np.array([])
    mask = np.asarray(mask)
    nz, ny, nx = mask.shape
    (xmin, xmax), (ymin, ymax), (zmin, zmax) = bounds
    eng = default_engine(device)
    m = torch.from_numpy(np.ascontiguousarray(mask != 0).view(np.uint8)).to(eng.device)
    lin = eng.boundary_voxels(m, thickness=int(thickness))
    if lin.numel() == 0:
        return np.array([]), np.array([]), np.array([])
    if sampling_step > 1:
        lin = lin[::sampling_step]
    lin = lin.cpu().numpy()
    Z_idx, rem = np.divmod(lin, ny * nx)
    Y_idx, X_idx = np.divmod(rem, nx)
    z_phys = zmin + Z_idx * (zmax - 1 - zmin) / (nz - 1) if nz > 1 else np.full_like(Z_idx, zmin)
    y_phys = ymin + Y_idx * (ymax - 1 - ymin) / (ny - 1) if ny > 1 else np.full_like(Y_idx, ymin)
    x_phys = xmin + X_idx * (xmax - 1 - xmin) / (nx - 1) if nx > 1 else np.full_like(X_idx, xmin)
    return x_phys, y_phys, z_phys
