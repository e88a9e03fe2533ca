"""ctypes binding of libptvb200.so (the C ABI declared in include/ptv_b200.h).

This is the stub a maintainer of the reference would add.  There is no CPU
fallback: if the shared library is missing and cannot be built, or no CUDA device is present,
every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PTV_OK, PTV_ERR_INVALID, PTV_ERR_TOO_FEW, PTV_ERR_CUDA, PTV_ERR_SINGULAR, PTV_ERR_NOMEM, PTV_ERR_QHULL = range(7)
METHOD_IDW, METHOD_SIBSON, METHOD_NEAREST, METHOD_RBF, METHOD_MADFILTER = range(5)
METHOD_RBF_CUBIC, METHOD_RBF_LINEAR, METHOD_RBF_QUINTIC = 5, 6, 7
METHOD_LINEAR = 8  # griddata(method='linear'): Delaunay tetrahedron + barycentric weights
F32, F64 = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libptvb200.so")
_lib = None

EXPORTS = [
    "ptv_version", "ptv_last_error", "ptv_device_info", "ptv_set_tuning", "ptv_get_tuning", "ptv_launch_count",
    "ptv_hash_create", "ptv_hash_destroy", "ptv_hash_build", "ptv_hash_build_slab", "ptv_hash_clip_violations", "ptv_hash_clip_violations_to", "ptv_hash_info", "ptv_knn_interp", "ptv_knn_stats", "ptv_knn_fail_reasons", "ptv_knn_work_stats", "ptv_linear_stats", "ptv_knn_points", "ptv_outlier_filter",
    "ptv_mask_gather", "ptv_boundary_voxels", "ptv_boundary_voxels_ws", "ptv_boundary_workspace_bytes", "ptv_apply_mask", "ptv_divergence", "ptv_divergence_flux",
    "ptv_flux_profiles", "ptv_strain_vorticity", "ptv_poisson_workspace_bytes", "ptv_poisson_lsqr", "ptv_projection_correct",
    "ptv_interpolate_host", "ptv_selftest_division", "ptv_strain_vorticity_slab",
]


class PTVError(RuntimeError):
    pass


def _declare(lib):
    vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
    lib.ptv_version.restype = i32
    lib.ptv_version.argtypes = []
    lib.ptv_last_error.restype = C.c_char_p
    lib.ptv_last_error.argtypes = []
    lib.ptv_device_info.restype = i32
    lib.ptv_device_info.argtypes = [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_size_t)]
    lib.ptv_set_tuning.restype = i32
    lib.ptv_set_tuning.argtypes = [C.c_char_p, f64]
    lib.ptv_get_tuning.restype = f64
    lib.ptv_get_tuning.argtypes = [C.c_char_p]
    lib.ptv_launch_count.restype = i64
    lib.ptv_launch_count.argtypes = []
    lib.ptv_hash_create.restype = i32
    lib.ptv_hash_create.argtypes = [C.POINTER(vp)]
    lib.ptv_hash_destroy.restype = i32
    lib.ptv_hash_destroy.argtypes = [vp]
    lib.ptv_hash_build.restype = i32
    lib.ptv_hash_build.argtypes = [vp, vp, vp, i64, f64, vp]
    lib.ptv_hash_build_slab.restype = i32
    lib.ptv_hash_build_slab.argtypes = [vp, vp, vp, i64, f64, f64, f64, i32, f64, vp]
    lib.ptv_hash_clip_violations.restype = i32
    lib.ptv_hash_clip_violations.argtypes = [vp, C.POINTER(i64)]
    lib.ptv_hash_clip_violations_to.restype = i32
    lib.ptv_hash_clip_violations_to.argtypes = [vp, vp, vp]
    lib.ptv_hash_info.restype = i32
    lib.ptv_hash_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i32 * 3), C.POINTER(f64 * 3), C.POINTER(f64),
                                  C.POINTER(i32)]
    lib.ptv_knn_interp.restype = i32
    lib.ptv_knn_interp.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, i32, i32, f64, f64, i32, vp, vp, vp, vp, vp, vp]
    lib.ptv_knn_points.restype = i32
    lib.ptv_knn_points.argtypes = [vp, vp, i32, i32, f64, f64, i32, vp, vp, vp, vp, vp, vp]
    lib.ptv_outlier_filter.restype = i32
    lib.ptv_outlier_filter.argtypes = [vp, i32, f64, vp, vp, vp]
    lib.ptv_knn_stats.restype = i32
    lib.ptv_knn_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    lib.ptv_knn_fail_reasons.restype = i32
    lib.ptv_knn_fail_reasons.argtypes = [vp, C.POINTER(i64 * 4)]
    lib.ptv_knn_work_stats.restype = i32
    lib.ptv_knn_work_stats.argtypes = [vp, C.POINTER(i64 * 8)]
    lib.ptv_linear_stats.restype = i32
    lib.ptv_linear_stats.argtypes = [vp, C.POINTER(i64 * 8)]
    lib.ptv_mask_gather.restype = i32
    lib.ptv_mask_gather.argtypes = [vp, i32, i32, i32, vp, i32, vp, i32, vp, i32, vp, vp]
    lib.ptv_boundary_voxels.restype = i32
    lib.ptv_boundary_voxels.argtypes = [vp, i32, i32, i32, i32, vp, i64, C.POINTER(i64), vp]
    lib.ptv_boundary_workspace_bytes.restype = i64
    lib.ptv_boundary_workspace_bytes.argtypes = [i32, i32, i32]
    lib.ptv_boundary_voxels_ws.restype = i32
    lib.ptv_boundary_voxels_ws.argtypes = [vp, i32, i32, i32, i32, vp, i32, vp, i64, C.POINTER(i64), vp]
    lib.ptv_apply_mask.restype = i32
    lib.ptv_apply_mask.argtypes = [vp, vp, vp, vp, i64, i32, vp]
    lib.ptv_divergence.restype = i32
    lib.ptv_divergence.argtypes = [vp, vp, vp, vp, i32, i32, i32, f64, f64, f64, vp, vp, vp, i32, vp, vp, vp]
    lib.ptv_divergence_flux.restype = i32
    lib.ptv_divergence_flux.argtypes = [vp, vp, vp, vp, i32, i32, i32, f64, f64, f64, vp, vp, vp, i32, vp, vp, vp, vp, vp,
                                        vp]
    lib.ptv_flux_profiles.restype = i32
    lib.ptv_flux_profiles.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.ptv_strain_vorticity_slab.restype = i32
    lib.ptv_strain_vorticity_slab.argtypes = [vp, vp, vp, vp, i32, i32, i32, f64, f64, f64, vp, vp, i32, vp, vp, vp]
    lib.ptv_selftest_division.restype = i32
    lib.ptv_selftest_division.argtypes = [f64, i64, C.c_uint64, C.POINTER(i64)]
    lib.ptv_strain_vorticity.restype = i32
    lib.ptv_strain_vorticity.argtypes = [vp, vp, vp, vp, i32, i32, i32, f64, f64, f64, i32, vp, vp, vp]
    lib.ptv_poisson_workspace_bytes.restype = i64
    lib.ptv_poisson_workspace_bytes.argtypes = [i32, i32, i32]
    lib.ptv_poisson_lsqr.restype = i32
    lib.ptv_poisson_lsqr.argtypes = [vp, i32, vp, i32, i32, i32, f64, f64, f64, f64, f64, f64, f64, i32, vp, vp,
                                     C.POINTER(f64 * 8), vp]
    lib.ptv_projection_correct.restype = i32
    lib.ptv_projection_correct.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, f64, f64, f64, i32, vp, vp, vp, vp]
    lib.ptv_interpolate_host.restype = i32
    lib.ptv_interpolate_host.argtypes = [vp, vp, i64, vp, i32, vp, i32, vp, i32, vp, i32, i32, f64, f64, i32,
                                         vp, vp, vp]


def load():
    """Load (building in-tree first if needed) the shared library; raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    # build() returns at once when the binary matches the sources (content hash next to the .so), so a
    # stale library with an older ABI is never loaded silently
    from . import build as _build
    _build.build()
    # PTV_LIB_PATH: load another build of the same ABI instead (A/B timing of kernel variants)
    lib = C.CDLL(os.environ.get("PTV_LIB_PATH") or LIB_PATH)
    _declare(lib)
    _lib = lib
    return lib


def check(rc: int):
    """Map a C status to the exception type the reference raises in the same situation."""
    if rc == PTV_OK:
        return
    msg = load().ptv_last_error().decode("utf-8", "replace")
    if rc == PTV_ERR_INVALID:
        raise ValueError(msg)
    if rc == PTV_ERR_TOO_FEW:
        raise IndexError(msg)
    if rc == PTV_ERR_SINGULAR:
        raise np.linalg.LinAlgError(msg)
    if rc == PTV_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == PTV_ERR_QHULL:
        try:  # what griddata raises from Qhull (a RuntimeError subclass)
            from scipy.spatial import QhullError
        except Exception:  # pragma: no cover
            QhullError = RuntimeError
        raise QhullError(msg)
    raise PTVError(msg)
