"""Device-resident engine: torch owns device memory and streams, libptvb200.so does the work.

The engine keeps one spatial hash (buffers reused across rebuilds, e.g. per PTV frame) and
exposes the hot path on device tensors; ``interpolator.py`` / ``physics.py`` wrap it with the
reference's NumPy signatures.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi

_METHODS = {"idw": _cabi.METHOD_IDW, "sibson": _cabi.METHOD_SIBSON, "nearest": _cabi.METHOD_NEAREST,
            "rbf": _cabi.METHOD_RBF, "linear": _cabi.METHOD_LINEAR}
# scipy.interpolate.RBFInterpolator kernels usable without `epsilon` (the reference never passes one)
_RBF_KERNELS = {"thin_plate_spline": _cabi.METHOD_RBF, "cubic": _cabi.METHOD_RBF_CUBIC,
                "linear": _cabi.METHOD_RBF_LINEAR, "quintic": _cabi.METHOD_RBF_QUINTIC}
_RBF_NEED_EPSILON = ("multiquadric", "inverse_multiquadric", "inverse_quadratic", "gaussian")


def method_code(method: str, rbf_kernel: str = "thin_plate_spline") -> int:
    """C ABI method code; mirrors RBFInterpolator's argument checks (_rbfinterp.py:276-289)."""
    if method not in _METHODS:
        raise NotImplementedError(f"method {method!r} is not on the CUDA path")
    if method != "rbf":
        return _METHODS[method]
    kern = str(rbf_kernel).lower()
    if kern in _RBF_KERNELS:
        return _RBF_KERNELS[kern]
    if kern in _RBF_NEED_EPSILON:
        raise ValueError("`epsilon` must be specified if `kernel` is not one of "
                         "{'linear', 'thin_plate_spline', 'cubic', 'quintic'}.")
    raise ValueError(f"`kernel` must be one of {sorted(list(_RBF_KERNELS) + list(_RBF_NEED_EPSILON))}.")


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _cabi.PTVError("ptv_interpolation_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise _cabi.PTVError(f"device must be a CUDA device, got {dev}")
    return dev


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return _cabi.F32
    if dt == torch.float64:
        return _cabi.F64
    raise ValueError(f"unsupported field dtype {dt}")


class PTVEngine:
    """One spatial hash + the kernels that consume it, bound to one CUDA device."""

    def __init__(self, device=None):
        self.lib = _cabi.load()
        self.device = _require_cuda(device)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_hash_create(C.byref(self._h)))
        self._keep = None  # tensors the hash borrows

    def close(self):
        if self._h:
            self.lib.ptv_hash_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ hash
    def build(self, points: torch.Tensor, values: torch.Tensor, cell_size: float = 0.0):
        """points, values: (Np,3) float64 CUDA tensors (df[['x','y','z']], df[['u','v','w']])."""
        if points.dtype != torch.float64 or values.dtype != torch.float64:
            raise ValueError("points and values must be float64")
        if points.ndim != 2 or points.shape[1] != 3 or values.shape != points.shape:
            raise ValueError("points and values must both have shape (Np, 3)")
        points = points.contiguous()
        values = values.contiguous()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_hash_build(self._h, _ptr(points), _ptr(values), points.shape[0],
                                                float(cell_size), self._stream()))
        self._keep = (points, values)
        self.n_particles = points.shape[0]

    def build_slab(self, points: torch.Tensor, values: torch.Tensor, z_lo: float, z_hi: float, k: int,
                   halo_factor: float = 2.5, cell_size: float = 0.0):
        """build() restricted to the particles a rank needs for the planes z_lo..z_hi: its slab plus a halo of
        ``halo_factor`` expected k-neighbour radii.  After interpolating, ``clip_violations()`` must be 0 --
        otherwise some search left the binned range and the frame has to be redone after a full build()."""
        if points.dtype != torch.float64 or values.dtype != torch.float64:
            raise ValueError("points and values must be float64")
        points = points.contiguous()
        values = values.contiguous()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_hash_build_slab(self._h, _ptr(points), _ptr(values), points.shape[0],
                                                     float(cell_size), float(z_lo), float(z_hi), int(k),
                                                     float(halo_factor), self._stream()))
        self._keep = (points, values)
        self.n_particles = points.shape[0]

    def clip_violations(self) -> int:
        """Voxels of the interpolate() calls since the last build whose k-th distance reached outside the
        z-range a slab hash binned (0 on a full hash).  Synchronises the device."""
        c = C.c_int64()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_hash_clip_violations(self._h, C.byref(c)))
        return int(c.value)

    def clip_violations_to(self, dst: torch.Tensor):
        """clip_violations() written into a one-element float64 CUDA tensor on the current stream (no sync)."""
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_hash_clip_violations_to(self._h, _ptr(dst), self._stream()))

    def hash_info(self):
        n = C.c_int64()
        dims = (C.c_int * 3)()
        origin = (C.c_double * 3)()
        cell = C.c_double()
        mx = C.c_int()
        _cabi.check(self.lib.ptv_hash_info(self._h, C.byref(n), C.byref(dims), C.byref(origin), C.byref(cell),
                                           C.byref(mx)))
        return {"n": n.value, "dims": tuple(dims), "origin": tuple(origin), "cell": cell.value,
                "max_cell_count": mx.value}

    # ------------------------------------------------------------------ interpolation
    def interpolate(self, ax_x, ax_y, ax_z, mask=None, method="idw", k=50, idw_power=2.0, smoothing=0.0,
                    out_dtype=torch.float32, out=None, return_knn=False, rbf_kernel="thin_plate_spline"):
        """Fused kNN + weights on the rectilinear grid ax_x (x) ax_y (x) ax_z (float64 CUDA axes).
        mask: optional (nz,ny,nx) uint8/bool CUDA tensor, non-zero = pore.  Returns a (3,nz,ny,nx)
        tensor (U,V,W) and, if ``return_knn``, (dist (nvox,k) float64, idx (nvox,k) int64)."""
        code = method_code(method, rbf_kernel)
        nx, ny, nz = ax_x.numel(), ax_y.numel(), ax_z.numel()
        for a in (ax_x, ax_y, ax_z):
            if a.dtype != torch.float64 or not a.is_cuda:
                raise ValueError("grid axes must be float64 CUDA tensors")
        ax_x, ax_y, ax_z = ax_x.contiguous(), ax_y.contiguous(), ax_z.contiguous()
        if mask is not None:
            if mask.dtype == torch.bool:
                mask = mask.view(torch.uint8)
            if mask.dtype != torch.uint8 or tuple(mask.shape) != (nz, ny, nx):
                raise ValueError("mask must be uint8/bool with shape (nz, ny, nx)")
            mask = mask.contiguous()
        if out is None:
            out = torch.empty((3, nz, ny, nx), dtype=out_dtype, device=self.device)
        elif tuple(out.shape) != (3, nz, ny, nx) or not out[0].is_contiguous():
            raise ValueError("out must be a (3, nz, ny, nx) tensor whose components are contiguous")
        if method == "nearest":
            k = 1
        if method == "linear":
            k = 4  # the lists hold the tetrahedron's vertex rows and barycentric weights
        kd = ki = None
        if return_knn:
            ki = torch.empty((nz * ny * nx, k), dtype=torch.int64, device=self.device)
            kd = torch.empty((nz * ny * nx, k), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_knn_interp(self._h, _ptr(ax_x), nx, _ptr(ax_y), ny, _ptr(ax_z), nz, _ptr(mask),
                                                code, int(k), float(idw_power), float(smoothing),
                                                _dtype_code(out.dtype), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]),
                                                _ptr(ki), _ptr(kd), self._stream()))
        if return_knn:
            return out, kd, ki
        return out

    def interpolate_to_host(self, ax_x, ax_y, ax_z, host_out, mask=None, dev_out=None, chunks=None, mask_host=None,
                            **kw):
        """interpolate() with the device->host copy of U,V,W overlapped with the search: the slab is
        processed in z-chunks on the current stream while a second stream drains finished chunks into
        ``host_out`` (a pinned (3,nz,ny,nx) CPU tensor) -- one copy per chunk and component, each a
        contiguous block.  Returns (device tensor, completion event): ``host_out`` may be read after
        ``event.synchronize()``.  A ``dev_out`` given by the caller may be reused for the next frame at
        once: the main stream waits for the previous drain before the first kernel overwrites it.
        ``mask_host`` (a C-contiguous bool/uint8 (nz,ny,nx) NumPy array in ordinary host memory) may be given
        instead of ``mask``: each z-chunk of it is staged to the device right before that chunk is searched,
        i.e. while the previous chunk's kernel runs."""
        nx, ny, nz = ax_x.numel(), ax_y.numel(), ax_z.numel()
        if dev_out is None:
            dev_out = torch.empty((3, nz, ny, nx), dtype=host_out.dtype, device=self.device)
        if tuple(host_out.shape) != (3, nz, ny, nx) or host_out.dtype != dev_out.dtype:
            raise ValueError("host_out must be a (3, nz, ny, nx) CPU tensor of the output dtype")
        if chunks is None:
            # chunks of >= 32 planes (one region layer of the streaming kernel); many small ones keep the
            # drain of the LAST chunk -- the part that cannot overlap anything -- short
            chunks = max(1, min(32, nz // 32))
        if not hasattr(self, "copy_stream"):
            self.copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        # a reused dev_out must not be overwritten while an earlier drain still reads it
        main.wait_stream(self.copy_stream)
        # the drain runs on copy_stream: keep the block out of the caching allocator until it is done
        dev_out.record_stream(self.copy_stream)
        cuts = [round(i * nz / chunks) for i in range(chunks + 1)]
        if mask_host is not None:
            from . import hostmem
            if tuple(mask_host.shape) != (nz, ny, nx):
                raise ValueError("mask must have shape (nz, ny, nx)")
            key = (nz, ny, nx)
            if getattr(self, "_mask_dev_key", None) != key:
                self._mask_dev = torch.empty(key, dtype=torch.uint8, device=self.device)
                self._mask_dev_key = key
            mask = self._mask_dev
            if not hasattr(self, "in_stream"):
                self.in_stream = torch.cuda.Stream(device=self.device)
            self.in_stream.wait_stream(main)  # earlier kernels may still read the device mask
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b <= a:
                continue
            if mask_host is not None:  # DMA on the input stream: it overlaps the previous chunk's kernel
                hostmem.stage_to_device(mask_host[a:b], self.device, out=mask[a:b], stream=self.in_stream)
                main.wait_stream(self.in_stream)
            self.interpolate(ax_x, ax_y, ax_z[a:b], mask=None if mask is None else mask[a:b],
                             out=dev_out[:, a:b], **kw)
            done = torch.cuda.Event()
            done.record(main)
            self.copy_stream.wait_event(done)
            with torch.cuda.stream(self.copy_stream):
                for c in range(3):
                    host_out[c, a:b].copy_(dev_out[c, a:b], non_blocking=True)
        finished = torch.cuda.Event()
        finished.record(self.copy_stream)
        return dev_out, finished

    def interpolate_points(self, queries: "PTVEngine", method="idw", k=50, idw_power=2.0, smoothing=0.0,
                           out_dtype=torch.float32, return_knn=False, values=True, rbf_kernel="thin_plate_spline"):
        """The same search / weights for arbitrary query points: ``queries`` is a second engine whose
        hash was built over the query points (pass ``self`` for a self-query).  Returns a (3, nq) tensor
        indexed by the query's original row (and (dist, idx) (nq,k) if ``return_knn``)."""
        code = method_code(method, rbf_kernel)
        nq = queries.n_particles
        if method == "nearest":
            k = 1
        if method == "linear":
            k = 4
        out = torch.empty((3, nq), dtype=out_dtype, device=self.device) if values else None
        kd = ki = None
        if return_knn:
            ki = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            kd = torch.empty((nq, k), dtype=torch.float64, device=self.device)
        o = [None, None, None] if out is None else [out[0], out[1], out[2]]
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_knn_points(self._h, queries._h, code, int(k), float(idw_power),
                                                float(smoothing), _dtype_code(out_dtype), _ptr(o[0]), _ptr(o[1]),
                                                _ptr(o[2]), _ptr(ki), _ptr(kd), self._stream()))
        if return_knn:
            return out, kd, ki
        return out

    def outlier_filter(self, k=25, threshold=3.0):
        """kNN median/MAD filter on the particles of the current hash (filtering.py:5-58).  Returns
        (keep uint8 (Np,), kth_dist float64 (Np,)) CUDA tensors."""
        n = self.n_particles
        keep = torch.empty(n, dtype=torch.uint8, device=self.device)
        kth = torch.empty(n, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_outlier_filter(self._h, int(k), float(threshold), _ptr(keep), _ptr(kth),
                                                    self._stream()))
        return keep, kth

    def knn_stats(self):
        """Diagnostics of the last interpolate(): did the streaming kernel run, how many tiles fell
        back to the exact heap kernel, how many it finished itself (needs set_tuning(stats=1))."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _cabi.check(self.lib.ptv_knn_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        r = (C.c_int64 * 4)()
        _cabi.check(self.lib.ptv_knn_fail_reasons(self._h, C.byref(r)))
        w = (C.c_int64 * 8)()
        _cabi.check(self.lib.ptv_knn_work_stats(self._h, C.byref(w)))
        keys = ("pairs_histogram", "pairs_classify", "exact_keys", "list_entries", "voxels", "rounds",
                "histogram_points", "retries")
        return {"used_stream": bool(a.value), "tiles_failed": b.value, "tiles_streamed": c.value,
                "fail_reasons": {"no_estimate": r[0], "beyond_range": r[1], "bin_overflow": r[2], "verify": r[3]},
                "work": dict(zip(keys, (int(v) for v in w)))}

    def linear_stats(self):
        """Diagnostics of the last method='linear' call (needs set_tuning(stats=1))."""
        r = (C.c_int64 * 8)()
        _cabi.check(self.lib.ptv_linear_stats(self._h, C.byref(r)))
        keys = ("shared_pass", "reused_tetrahedra", "general_path", "global_candidate_sets", "outside_hull",
                "unresolved", "pivots", "hull_candidates")
        return dict(zip(keys, (int(v) for v in r)))

    # ------------------------------------------------------------------ grid ops
    def mask_gather(self, mask_raw, ix, iy, iz):
        rnz, rny, rnx = mask_raw.shape
        if mask_raw.dtype == torch.bool:
            mask_raw = mask_raw.view(torch.uint8)
        mask_raw = mask_raw.contiguous()
        out = torch.empty((iz.numel(), iy.numel(), ix.numel()), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_mask_gather(_ptr(mask_raw), rnx, rny, rnz, _ptr(ix), ix.numel(), _ptr(iy),
                                                 iy.numel(), _ptr(iz), iz.numel(), _ptr(out), self._stream()))
        return out

    def boundary_voxels(self, mask, thickness=1):
        """Linear C-order indices (int64 CUDA tensor) of solid voxels within ``thickness``
        6-connected dilation steps of fluid."""
        nz, ny, nx = mask.shape
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        mask = mask.contiguous()
        cnt = C.c_int64()
        nbytes = int(self.lib.ptv_boundary_workspace_bytes(nx, ny, nz))
        if getattr(self, "_bnd_work", None) is None or self._bnd_work.numel() < nbytes:
            self._bnd_work = torch.empty(nbytes, dtype=torch.uint8, device=self.device)  # kept for the next call
        work = self._bnd_work
        with torch.cuda.device(self.device):
            # phase 0: bit-pack, dilate, flag, count; phase 1: ordered index write from the flags left in `work`
            _cabi.check(self.lib.ptv_boundary_voxels_ws(_ptr(mask), nx, ny, nz, int(thickness), _ptr(work), 0, None, 0,
                                                        C.byref(cnt), self._stream()))
            idx = torch.empty((cnt.value,), dtype=torch.int64, device=self.device)
            if cnt.value:
                _cabi.check(self.lib.ptv_boundary_voxels_ws(_ptr(mask), nx, ny, nz, int(thickness), _ptr(work), 1,
                                                            _ptr(idx), cnt.value, C.byref(cnt), self._stream()))
        return idx

    def apply_mask(self, uvw, mask):
        if mask is not None and mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        n = uvw[0].numel()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_apply_mask(_ptr(uvw[0]), _ptr(uvw[1]), _ptr(uvw[2]), _ptr(mask), n,
                                                _dtype_code(uvw.dtype), self._stream()))
        return uvw

    def divergence(self, u, v, w, mask, dx, dy, dz, w_below=None, w_above=None, mask_above=None,
                   with_stats=False):
        nz, ny, nx = u.shape
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        if mask_above is not None and mask_above.dtype == torch.bool:
            mask_above = mask_above.view(torch.uint8)
        u, v, w, mask = u.contiguous(), v.contiguous(), w.contiguous(), mask.contiguous()
        div = torch.empty_like(u)
        stats = torch.zeros(2, dtype=torch.float64, device=self.device) if with_stats else None
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_divergence(_ptr(u), _ptr(v), _ptr(w), _ptr(mask), nx, ny, nz, float(dx),
                                                float(dy), float(dz), _ptr(w_below), _ptr(w_above),
                                                _ptr(mask_above), _dtype_code(u.dtype), _ptr(div), _ptr(stats),
                                                self._stream()))
        return (div, stats) if with_stats else div

    def divergence_flux(self, u, v, w, mask, dx, dy, dz, w_below=None, w_above=None, mask_above=None, z0=0,
                        nz_global=None, extra=0):
        """One pass: (div, stats[2] = (sum|div| over fluid, n_fluid), q_xy[nz], q_xz[ny], q_yz[nx]).
        With ``nz_global`` the accumulators live in one flat buffer laid out for the whole grid -- Q_xy of
        this slab at planes z0.. -- so that a single all-reduce finishes all of them (returned as the sixth
        value; q_xy is then the full-length profile; ``extra`` more zeroed slots follow Q_yz for the caller)."""
        nz, ny, nx = u.shape
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        if mask_above is not None and mask_above.dtype == torch.bool:
            mask_above = mask_above.view(torch.uint8)
        u, v, w, mask = u.contiguous(), v.contiguous(), w.contiguous(), mask.contiguous()
        div = torch.empty_like(u)
        nzg = nz if nz_global is None else int(nz_global)
        acc = torch.zeros(2 + nzg + ny + nx + extra, dtype=torch.float64, device=self.device)
        stats, qxy, qxz, qyz = (acc[:2], acc[2:2 + nzg], acc[2 + nzg:2 + nzg + ny],
                                acc[2 + nzg + ny:2 + nzg + ny + nx])
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_divergence_flux(_ptr(u), _ptr(v), _ptr(w), _ptr(mask), nx, ny, nz, float(dx),
                                                     float(dy), float(dz), _ptr(w_below), _ptr(w_above),
                                                     _ptr(mask_above), _dtype_code(u.dtype), _ptr(div), _ptr(stats),
                                                     _ptr(qxy[z0:z0 + nz]), _ptr(qxz), _ptr(qyz), self._stream()))
        if nz_global is None:
            return div, stats, qxy, qxz, qyz
        return div, stats, qxy, qxz, qyz, acc

    def strain_vorticity(self, u, v, w, dx, dy, dz, mask=None, strain=True, vorticity=True, below=None, above=None):
        """(shear-rate magnitude, vorticity magnitude) of the field; either may be skipped (None).
        ``below`` / ``above``: for a z-slab of a sharded grid, the (3, ny, nx) planes (u, v, w) of the z-neighbours
        just outside the slab (``SlabComm.exchange_planes``); None = a face of the whole domain."""
        nz, ny, nx = u.shape
        if mask is not None and mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        u, v, w = u.contiguous(), v.contiguous(), w.contiguous()
        if below is not None or above is not None:
            return self._strain_vorticity_slab(u, v, w, dx, dy, dz, mask, strain, vorticity, below, above)
        s = torch.empty_like(u) if strain else None
        o = torch.empty_like(u) if vorticity else None
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_strain_vorticity(_ptr(u), _ptr(v), _ptr(w), _ptr(mask), nx, ny, nz, float(dx),
                                                      float(dy), float(dz), _dtype_code(u.dtype), _ptr(s), _ptr(o),
                                                      self._stream()))
        return s, o

    def _strain_vorticity_slab(self, u, v, w, dx, dy, dz, mask, strain, vorticity, below, above):
        nz, ny, nx = u.shape
        for h in (below, above):
            if h is not None and (tuple(h.shape) != (3, ny, nx) or h.dtype != u.dtype):
                raise ValueError("halo planes must be (3, ny, nx) tensors of the fields' dtype")
        below = None if below is None else below.contiguous()
        above = None if above is None else above.contiguous()
        s = torch.empty_like(u) if strain else None
        o = torch.empty_like(u) if vorticity else None
        if u.dtype == torch.float32 and nx % 16 == 0:
            with torch.cuda.device(self.device):
                rc = self.lib.ptv_strain_vorticity_slab(_ptr(u), _ptr(v), _ptr(w), _ptr(mask), nx, ny, nz, float(dx),
                                                        float(dy), float(dz), _ptr(below), _ptr(above),
                                                        _dtype_code(u.dtype), _ptr(s), _ptr(o), self._stream())
            if rc == 0:
                return s, o
            if rc != 1:  # anything but "this shape / alignment is not served" (PTV_ERR_INVALID) is a real error
                _cabi.check(rc)
        # shapes the slab kernel does not take: the slab padded with its halo planes through the whole-grid kernels;
        # the padded planes' own results (one-sided differences) are dropped
        lo, hi = (0 if below is None else 1), (0 if above is None else 1)

        def pad(f, i):
            parts = ([below[i:i + 1]] if lo else []) + [f] + ([above[i:i + 1]] if hi else [])
            return torch.cat(parts, dim=0)

        mp = None
        if mask is not None:
            ones = torch.ones((1, ny, nx), dtype=mask.dtype, device=mask.device)
            mp = torch.cat(([ones] if lo else []) + [mask] + ([ones] if hi else []), dim=0)
        sp, op = self.strain_vorticity(pad(u, 0), pad(v, 1), pad(w, 2), dx, dy, dz, mask=mp, strain=strain,
                                       vorticity=vorticity)
        cut = slice(lo, lo + nz)
        return (None if sp is None else sp[cut].contiguous()), (None if op is None else op[cut].contiguous())

    def poisson_lsqr(self, div, mask, dx, dy, dz, damp=1e-8, atol=1e-10, btol=1e-10, conlim=1e8, iter_lim=3000):
        """LSQR solve of the masked Laplacian system for div - mean(div[mask]) (physics.py:180-186).
        Returns (phi float64 (nz,ny,nx), info dict)."""
        nz, ny, nx = div.shape
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        div, mask = div.contiguous(), mask.contiguous()
        phi = torch.empty((nz, ny, nx), dtype=torch.float64, device=self.device)
        nbytes = int(self.lib.ptv_poisson_workspace_bytes(nx, ny, nz))
        work = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=self.device)
        info = (C.c_double * 8)()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_poisson_lsqr(_ptr(div), _dtype_code(div.dtype), _ptr(mask), nx, ny, nz, float(dx),
                                                  float(dy), float(dz), float(damp), float(atol), float(btol),
                                                  float(conlim), int(iter_lim), _ptr(phi), _ptr(work), C.byref(info),
                                                  self._stream()))
        keys = ("istop", "itn", "r1norm", "r2norm", "anorm", "acond", "arnorm", "xnorm")
        return phi, {k: (int(v) if k in ("istop", "itn") else float(v)) for k, v in zip(keys, info)}

    def projection_correct(self, u, v, w, phi, mask, dx, dy, dz):
        """apply_consistent_correction (physics.py:110-147): returns the corrected (u, v, w)."""
        nz, ny, nx = u.shape
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        u, v, w, mask, phi = u.contiguous(), v.contiguous(), w.contiguous(), mask.contiguous(), phi.contiguous()
        out = torch.empty((3, nz, ny, nx), dtype=u.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_projection_correct(_ptr(u), _ptr(v), _ptr(w), _ptr(phi), _ptr(mask), nx, ny, nz,
                                                        float(dx), float(dy), float(dz), _dtype_code(u.dtype),
                                                        _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), self._stream()))
        return out[0], out[1], out[2]

    def flux_profiles(self, u, v, w):
        """Unscaled plane sums (q_xy[nz], q_xz[ny], q_yz[nx]) as float64 CUDA tensors."""
        ref = next(t for t in (u, v, w) if t is not None)
        nz, ny, nx = ref.shape
        qxy = torch.zeros(nz, dtype=torch.float64, device=self.device)
        qxz = torch.zeros(ny, dtype=torch.float64, device=self.device)
        qyz = torch.zeros(nx, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ptv_flux_profiles(_ptr(u), _ptr(v), _ptr(w), nx, ny, nz, _dtype_code(ref.dtype),
                                                   _ptr(qxy), _ptr(qxz), _ptr(qyz), self._stream()))
        return qxy, qxz, qyz


_default_engine = {}


def default_engine(device=None) -> PTVEngine:
    dev = _require_cuda(device)
    key = (dev.type, dev.index)
    if key not in _default_engine:
        _default_engine[key] = PTVEngine(dev)
    return _default_engine[key]


def set_tuning(**kw):
    lib = _cabi.load()
    for k, v in kw.items():
        _cabi.check(lib.ptv_set_tuning(k.encode(), float(v)))
