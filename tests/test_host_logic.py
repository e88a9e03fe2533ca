"""CPU-only tests: the C ABI library loads and exports every declared symbol, and the host-side
logic of the drop-in modules matches the oracle.  No compute entry point is called here."""
import os
import re

import numpy as np
import pytest

from oracle import reference_port as rp
from ptv_interpolation_b200 import _cabi
from ptv_interpolation_b200 import interpolator as gi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ptv_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ptv_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_cabi.EXPORTS) == declared
    assert lib.ptv_version() >= 100
    assert lib.ptv_set_tuning(b"no_such_knob", 1.0) == _cabi.PTV_ERR_INVALID
    assert b"unknown key" in lib.ptv_last_error()
    assert lib.ptv_get_tuning(b"tile") == 128


def test_error_mapping():
    lib = _cabi.load()
    lib.ptv_set_tuning(b"bogus", 0.0)
    with pytest.raises(ValueError):
        _cabi.check(_cabi.PTV_ERR_INVALID)
    with pytest.raises(IndexError):
        _cabi.check(_cabi.PTV_ERR_TOO_FEW)
    with pytest.raises(np.linalg.LinAlgError):
        _cabi.check(_cabi.PTV_ERR_SINGULAR)
    with pytest.raises(_cabi.PTVError):
        _cabi.check(_cabi.PTV_ERR_CUDA)
    from scipy.spatial import QhullError
    with pytest.raises(QhullError):  # method='linear': what griddata raises from Qhull
        _cabi.check(_cabi.PTV_ERR_QHULL)
    # the numeric codes are the header's
    hdr = open(os.path.join(ROOT, "include", "ptv_b200.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define (PTV_[A-Z0-9_]+) (\d+)", hdr)}
    assert defs["PTV_METHOD_LINEAR"] == _cabi.METHOD_LINEAR and defs["PTV_ERR_QHULL"] == _cabi.PTV_ERR_QHULL
    assert defs["PTV_METHOD_IDW"] == _cabi.METHOD_IDW and defs["PTV_METHOD_RBF_QUINTIC"] == _cabi.METHOD_RBF_QUINTIC


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pandas as pd
    df = pd.DataFrame({c: np.arange(60.0) for c in "xyzuvw"})
    grid, _ = gi.create_grid(((0, 4), (0, 4), (0, 4)), 4)
    with pytest.raises(_cabi.PTVError):
        gi.interpolate_field(df, grid, method="idw", idw_neighbors=5)
    with pytest.raises(_cabi.PTVError):  # the reference's default method has no CPU path either
        gi.interpolate_field(df, grid)
    with pytest.raises(ValueError):  # griddata's own error for 3-D 'cubic' (interpolator.py:197)
        gi.interpolate_field(df, grid, method="cubic")
    from ptv_interpolation_b200 import physics
    with pytest.raises(_cabi.PTVError):
        physics.compute_consistent_divergence(*(np.zeros((3, 3, 3)),) * 3, np.ones((3, 3, 3), bool), 1, 1, 1)


@pytest.mark.parametrize("bounds,res", [(((0, 14), (0, 12), (0, 10)), (13, 11, 9)),
                                        (((3, 19), (-2, 16), (1.5, 21.5)), 7),
                                        (((0, 2048), (0, 2048), (0, 2048)), (5, 4, 3))])
def test_create_grid_matches_reference_port(bounds, res):
    (X, Y, Z), (x, y, z) = gi.create_grid(bounds, res)
    (Xr, Yr, Zr), (xr, yr, zr) = rp.create_grid(bounds, res)
    assert np.array_equal(x, xr) and np.array_equal(y, yr) and np.array_equal(z, zr)
    assert np.array_equal(X, Xr) and np.array_equal(Y, Yr) and np.array_equal(Z, Zr)
    ax = gi._grid_axes((X, Y, Z))
    assert all(np.array_equal(a, b) for a, b in zip(ax, (x, y, z)))
    ax = gi._grid_axes((Xr, Yr, Zr))  # dense meshgrid is verified, not assumed
    assert all(np.array_equal(a, b) for a, b in zip(ax, (x, y, z)))
    # anything else is treated as arbitrary query points (point-query kernel)
    assert gi._grid_axes((Xr + np.random.default_rng(0).random(Xr.shape), Yr, Zr)) is None


def test_nearest_axis_index_matches_oracle():
    rng = np.random.default_rng(1)
    g = np.linspace(2.0, 17.0, 9)
    q = np.concatenate([rng.uniform(0, 20, 500), g, (g[:-1] + g[1:]) / 2, [2.0, 17.0, 1.999, 17.001]])
    assert np.array_equal(gi._nearest_axis_index(g, q), rp.nearest_axis_index(g, q))
    assert np.array_equal(gi._nearest_axis_index(np.array([3.0]), np.array([3.0, 4.0])), [0, -1])


def test_load_errors_are_ioerror(tmp_path):
    with pytest.raises(IOError):
        gi.load_ptv_data(str(tmp_path / "missing.csv"))
    p = tmp_path / "bad.csv"
    p.write_text("a,b\n1,2\n")
    with pytest.raises(IOError):
        gi.load_ptv_data(str(p))
    p2 = tmp_path / "ok.csv"
    p2.write_text("x,y,z,vx,vy,vz\n1,2,3,4,5,6\n")
    assert list(gi.load_ptv_data(str(p2)).columns) == ["x", "y", "z", "u", "v", "w"]
    with pytest.raises(IOError):
        gi.load_mask(str(tmp_path / "missing.tif"))


def test_synthetic_generators_small():
    from ptv_interpolation_b200 import synthetic
    m = synthetic.hex6_sphere_pack_mask(32)
    assert m.shape == (32, 32, 32) and 0.3 < m.float().mean() < 0.9
    f = synthetic.fcc_sphere_pack_mask(48, lattice=24.0)
    assert abs(float(f.float().mean()) - 0.40) < 0.03
    slab = synthetic.fcc_sphere_pack_mask(48, lattice=24.0, z0=10, nz_local=5)
    assert np.array_equal(slab.numpy(), f[10:15].numpy())
    p = synthetic.sample_pore_particles(m, 1000, seed=3)
    assert p.shape == (1000, 3) and p.dtype.is_floating_point
    pn = p.numpy()
    assert np.array_equal(pn, pn.astype(np.float32).astype(np.float64))
    ix, iy, iz = (np.rint(pn[:, c]).astype(int) for c in range(3))
    assert m.numpy()[iz, iy, ix].all()
    assert np.array_equal(p.numpy(), synthetic.sample_pore_particles(m, 1000, seed=3).numpy())
    c = synthetic.cylinder_array_mask(64)
    assert 0.55 < float(c.float().mean()) < 0.68
    v = synthetic.cylinder_flow(p)
    assert v.shape == (1000, 3) and np.isfinite(v.numpy()).all()


def test_compat_aliases_modules(tmp_path):
    """compat.install(): the reference's import names resolve to the CUDA drop-ins, other names fall
    through to the script directory's own modules."""
    import importlib.util
    import sys
    from ptv_interpolation_b200 import compat
    (tmp_path / "physics.py").write_text("def solve_poisson():\n    return 'reference'\n"
                                         "def compute_consistent_divergence():\n    return 'reference'\n")
    saved = {k: sys.modules.get(k) for k in ("interpolator", "physics", "filtering", "velocity_analysis")}
    try:
        compat.install(str(tmp_path))
        import interpolator, physics, filtering  # noqa: E401
        from ptv_interpolation_b200 import interpolator as gi2, physics as gp2
        assert interpolator.interpolate_field is gi2.interpolate_field
        assert physics.compute_consistent_divergence is gp2.compute_consistent_divergence
        assert physics.solve_poisson() == "reference"
        assert hasattr(filtering, "apply_filters")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
