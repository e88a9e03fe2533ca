"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the
golden vectors of the live reference.  Bars: neighbour indices, distances and mask bits exact;
velocities |gpu-ref| <= 1e-5 * max(|ref|, rms(input component))."""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import reference_port as rp  # noqa: E402
from ptv_interpolation_b200 import interpolator as gi  # noqa: E402
from ptv_interpolation_b200 import physics as gp  # noqa: E402
from ptv_interpolation_b200 import synthetic  # noqa: E402
from ptv_interpolation_b200.engine import PTVEngine, set_tuning  # noqa: E402

TOL = 1e-5


def _df(points, values):
    return pd.DataFrame({"x": points[:, 0], "y": points[:, 1], "z": points[:, 2],
                         "u": values[:, 0], "v": values[:, 1], "w": values[:, 2]})


def _bounds(b):
    return tuple(tuple(float(v) for v in row) for row in b)


def _assert_vel(got, ref, values, tol=TOL):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    for c in range(3):
        s = np.sqrt(np.mean(values[:, c] ** 2))
        bound = tol * np.maximum(np.abs(ref[c]), s)
        err = np.abs(got[c] - ref[c])
        assert np.all(err <= bound), (c, float(err.max()), float((err / bound).max()))


@pytest.fixture(scope="module")
def case_a(golden_dir):
    g = np.load(os.path.join(golden_dir, "case_a_interp.npz"))
    grid, axes = gi.create_grid(_bounds(g["bounds"]), tuple(int(r) for r in g["res"]))
    return g, grid, _df(g["points"], g["values"])


@pytest.mark.parametrize("name,kw", [
    ("idw_k50", dict(method="idw")),
    ("idw_k8_p3", dict(method="idw", idw_neighbors=8, idw_power=3.0)),
    ("idw_k8_p15", dict(method="idw", idw_neighbors=8, idw_power=1.5)),
    ("sibson_k30", dict(method="sibson")),
    ("sibson_k12", dict(method="sibson", sibson_neighbors=12)),
    ("nearest", dict(method="nearest")),
])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_golden_interp(case_a, name, kw, dtype):
    g, grid, df = case_a
    U, V, W = gi.interpolate_field(df, grid, out_dtype=dtype, **kw)
    assert U.dtype == dtype and U.shape == grid[0].shape and U.flags.writeable
    _assert_vel(np.stack([U, V, W]), g[name], g["values"])
    if dtype == np.float64:  # fp64 accumulate: far inside the fp32 bar
        _assert_vel(np.stack([U, V, W]), g[name], g["values"], tol=1e-11)


@pytest.mark.parametrize("k", [1, 8, 50])
def test_golden_knn_bitexact(case_a, k):
    g, grid, df = case_a
    method = "nearest" if k == 1 else "idw"
    U, V, W, kd, ki = gi.interpolate_field(df, grid, method=method, idw_neighbors=k, return_knn=True)
    assert np.array_equal(ki, g[f"knn_i_k{k}"])
    assert np.array_equal(kd, g[f"knn_d_k{k}"])


def test_lattice_ties_canonical(golden_dir):
    g = np.load(os.path.join(golden_dir, "case_b_boundary.npz"))
    b = _bounds(g["bounds"])
    grid, _ = gi.create_grid(b, 12)
    pts, vals = g["points"], g["values"]
    U, V, W, kd, ki = gi.interpolate_field(_df(pts, vals), grid, method="idw", idw_neighbors=20,
                                           return_knn=True, out_dtype=np.float64)
    _assert_vel(np.stack([U, V, W]), g["idw_k20"], vals)
    og, _ = rp.create_grid(b, 12)
    d, i, _ = rp.knn_canonical(pts, rp.flat_coords(og), 20)
    assert np.array_equal(ki, i)
    assert np.array_equal(kd, d)
    # the same through the brute-force definition
    d2, i2, _ = rp.knn_bruteforce(pts, rp.flat_coords(og), 20)
    assert np.array_equal(ki, i2) and np.array_equal(kd, d2)


def test_boundary_particles_and_mask_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "case_b_boundary.npz"))
    b = _bounds(g["bounds"])
    for th, st in ((1, 1), (2, 1), (2, 3), (3, 5)):
        bx, by, bz = gi.extract_boundary_particles(g["mask_raw"], b, sampling_step=st, thickness=th)
        assert np.array_equal(np.stack([bx, by, bz], 0).astype(np.float64), g[f"bp_t{th}_s{st}"])
    grid, _ = gi.create_grid(b, 12)
    assert np.array_equal(gi.sample_mask_on_grid(g["mask_raw"], grid, b), g["mask_grid"])
    none = gi.extract_boundary_particles(np.ones((4, 5, 6), bool), ((0, 6), (0, 5), (0, 4)))
    assert all(len(a) == 0 for a in none)


@pytest.mark.parametrize("name", ["same", "down2", "down3", "shift", "up"])
def test_mask_sampling_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "case_c_mask.npz"))
    grid, _ = gi.create_grid(_bounds(g[f"{name}_bgrid"]), tuple(int(r) for r in g[f"{name}_res"]))
    out = gi.sample_mask_on_grid(g["mask_raw"], grid, _bounds(g[f"{name}_braw"]))
    assert out.dtype == bool and np.array_equal(out, g[f"{name}_out"])


def test_divergence_flux_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "case_d_divergence.npz"))
    dx, dy, dz = (float(h) for h in g["h"])
    div = gp.compute_consistent_divergence(g["u"], g["v"], g["w"], g["mask"], dx, dy, dz)
    assert div.dtype == np.float64 and np.array_equal(div, g["div"])
    assert np.allclose(gp.calculate_flux_xy(g["w"], dx, dy), g["q_xy"], rtol=1e-12, atol=1e-12)
    assert np.allclose(gp.calculate_flux_xz(g["v"], dx, dz), g["q_xz"], rtol=1e-12, atol=1e-12)
    assert np.allclose(gp.calculate_flux_yz(g["u"], dy, dz), g["q_yz"], rtol=1e-12, atol=1e-12)
    assert abs(gp.mid_plane_x_flux(g["u"], dy, dz) - float(g["mid_x"])) <= 1e-12 * max(1, abs(float(g["mid_x"])))
    m = gp.mean_abs_divergence(g["u"], g["v"], g["w"], g["mask"], dx, dy, dz)
    assert abs(m - float(g["mean_abs_div"])) <= 1e-12 * float(g["mean_abs_div"])
    # float32 fields: stencil evaluated in float64 on the upcast field, result rounded to float32
    u32, v32, w32 = (g[c].astype(np.float32) for c in "uvw")
    d32 = gp.compute_consistent_divergence(u32, v32, w32, g["mask"], dx, dy, dz)
    ref = rp.compute_consistent_divergence(u32.astype(np.float64), v32.astype(np.float64),
                                           w32.astype(np.float64), g["mask"], dx, dy, dz)
    assert d32.dtype == np.float32 and np.array_equal(d32, ref.astype(np.float32))


def test_exact_division_by_spacing():
    """The stencil kernels divide by the grid spacing with a reciprocal + two FMA corrections; every result must
    be the IEEE quotient (bit-identity with physics.py:26-53 / np.gradient rests on it)."""
    import ctypes as C
    from ptv_interpolation_b200 import _cabi
    lib = _cabi.load()
    for h in (2.00625, 4.0125, 1.25, 0.75, 3.0, 0.1, 1e-3, 7.0, -2.00625, 1.9999999999999998, 1.0000000000000002,
              1.0 / 3.0, 123456.789):
        bad = C.c_int64(-1)
        _cabi.check(lib.ptv_selftest_division(h, 1 << 28, 12345, C.byref(bad)))
        assert bad.value == 0, h


# nx % 16 == 0 -> bulk-async row pipeline (one chunk, two chunks with x halos, ragged last chunk, > 256 rows per
# z-plane so a plane is swept by several CTAs); the other shapes take the direct-load kernels
@pytest.mark.parametrize("shape", [(11, 9, 7), (6, 10, 16), (5, 33, 1028), (3, 4, 2052), (4, 37, 1040), (6, 5, 2080),
                                   (7, 300, 48), (2, 3, 1024), (1, 1, 16), (1, 5, 32), (5, 1, 64), (2, 66, 16)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_divergence_flux(shape, dtype):
    rng = np.random.default_rng(shape[2])
    u, v, w = (rng.normal(size=shape).astype(dtype) for _ in range(3))
    m = rng.random(shape) > 0.35
    h = (1.25, 0.75, 2.00625) if shape[0] > 5 else (1.0, 1.0, 1.0)
    eng = PTVEngine()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    div, st, qxy, qxz, qyz = eng.divergence_flux(t(u), t(v), t(w), t(m), *h)
    u64, v64, w64 = (a.astype(np.float64) for a in (u, v, w))
    ref = rp.compute_consistent_divergence(u64, v64, w64, m, *h)
    assert np.array_equal(div.cpu().numpy(), ref.astype(dtype))
    s, c = st.cpu().numpy()
    assert c == m.sum()
    assert abs(s / c - np.mean(np.abs(ref.astype(dtype).astype(np.float64)[m]))) <= 1e-12
    assert np.allclose(qxy.cpu().numpy(), w64.sum(axis=(1, 2)), rtol=1e-12, atol=1e-10)
    assert np.allclose(qxz.cpu().numpy(), v64.sum(axis=(0, 2)), rtol=1e-12, atol=1e-10)
    assert np.allclose(qyz.cpu().numpy(), u64.sum(axis=(0, 1)), rtol=1e-12, atol=1e-10)
    # slab form with halos reproduces the whole-grid result
    if shape[0] >= 5:
        a, b = 2, 4
        d2, st2, *_ = eng.divergence_flux(t(u[a:b]), t(v[a:b]), t(w[a:b]), t(m[a:b]), *h, w_below=t(w[a - 1]),
                                          w_above=t(w[b]), mask_above=t(m[b]).view(torch.uint8))
        assert np.array_equal(d2.cpu().numpy(), ref[a:b].astype(dtype))


def test_divergence_slab_halos_match_whole():
    rng = np.random.default_rng(9)
    shape = (12, 10, 9)
    u, v, w = (rng.normal(size=shape) for _ in range(3))
    m = rng.random(shape) > 0.3
    ref = rp.compute_consistent_divergence(u, v, w, m, 1.0, 2.0, 0.5)
    eng = PTVEngine()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pieces, tot = [], torch.zeros(2, dtype=torch.float64, device="cuda")
    cuts = [0, 5, 6, 12]
    for a, b in zip(cuts[:-1], cuts[1:]):
        below = t(w[a - 1]) if a > 0 else None
        above = t(w[b]) if b < shape[0] else None
        mabove = t(m[b]).view(torch.uint8) if b < shape[0] else None
        d, st = eng.divergence(t(u[a:b]), t(v[a:b]), t(w[a:b]), t(m[a:b]), 1.0, 2.0, 0.5, w_below=below,
                               w_above=above, mask_above=mabove, with_stats=True)
        pieces.append(d.cpu().numpy())
        tot += st
    assert np.array_equal(np.concatenate(pieces, 0), ref)
    s, c = tot.cpu().numpy()
    assert c == m.sum() and abs(s / c - rp.mean_abs_div(ref, m)) < 1e-13


@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("method,k", [("idw", 50), ("sibson", 30), ("idw", 3)])
def test_sphere_pack_vs_oracle(masked, method, k):
    n = 40
    mask = synthetic.hex6_sphere_pack_mask(n)
    pts = synthetic.sample_pore_particles(mask, 6000, seed=11)
    vals = synthetic.sphere_pack_flow(pts, n)
    pts, vals, mask = pts.numpy(), vals.numpy(), mask.numpy()
    b = ((0, n), (0, n), (0, n))
    grid, _ = gi.create_grid(b, n)
    kw = dict(method=method, idw_neighbors=k, sibson_neighbors=k)
    U, V, W, kd, ki = gi.interpolate_field(_df(pts, vals), grid, mask=mask if masked else None,
                                           return_knn=True, **kw)
    from ptv_interpolation_b200.engine import default_engine
    # k >= 8: the neighbour lists come from the PRODUCTION streaming kernel (knn_duo_kernel + canonical sort)
    assert default_engine().knn_stats()["used_stream"] == (k >= 8)
    og, _ = rp.create_grid(b, n)
    Ur, Vr, Wr, d, i = rp.interpolate_field(pts, vals, og, return_knn=True, **kw)
    ref = np.stack([Ur, Vr, Wr])
    if masked:
        ref = np.stack(rp.apply_mask_zero(Ur, Vr, Wr, mask))
        sel = mask.ravel()
        assert np.array_equal(ki[sel], i[sel]) and np.array_equal(kd[sel], d[sel])
        assert np.all(ki[~sel] == -1)
        assert np.all(np.stack([U, V, W])[:, ~mask] == 0)
    else:
        assert np.array_equal(ki, i) and np.array_equal(kd, d)
    _assert_vel(np.stack([U, V, W]), ref, vals)


@pytest.mark.parametrize("tune", [dict(tile=64), dict(tile=32), dict(r0=0), dict(r0=3), dict(ppc=0.3),
                                  dict(ppc=6.0)])
def test_tuning_does_not_change_neighbours(case_a, tune):
    g, grid, df = case_a
    try:
        set_tuning(**tune)
        U, V, W, kd, ki = gi.interpolate_field(df, grid, method="idw", idw_neighbors=50, return_knn=True)
    finally:
        set_tuning(tile=128, r0=1, ppc=0.5)
    assert np.array_equal(ki, g["knn_i_k50"]) and np.array_equal(kd, g["knn_d_k50"])


STREAM_DEFAULT = 2  # Tuning::stream: 2 = warp-private streaming kernel (production), 1 = CTA-wide one


def _stream_vs_heap(df, grid, mask=None, stream=STREAM_DEFAULT, **kw):
    """float64 outputs of a streaming kernel and of the exact heap kernel: the neighbour SETS are
    identical iff the weighted means agree to summation-order rounding."""
    from ptv_interpolation_b200.engine import default_engine
    try:
        set_tuning(stream=stream, stats=1)
        a = np.stack(gi.interpolate_field(df, grid, mask=mask, out_dtype=np.float64, **kw))
        st = default_engine().knn_stats()
        set_tuning(stream=0)
        b = np.stack(gi.interpolate_field(df, grid, mask=mask, out_dtype=np.float64, **kw))
        assert not default_engine().knn_stats()["used_stream"]
    finally:
        set_tuning(stream=STREAM_DEFAULT, stats=0)
    assert st["used_stream"]
    scale = np.abs(b).max()
    assert np.abs(a - b).max() <= 1e-11 * scale, float(np.abs(a - b).max() / scale)
    return a, st


@pytest.mark.parametrize("tune", [dict(), dict(stream_tile=64), dict(stream_tile=32), dict(r0=0), dict(r0=2),
                                  dict(ppc=0.4), dict(ppc=5.0)])
@pytest.mark.parametrize("kw", [dict(method="idw"), dict(method="idw", idw_neighbors=9, idw_power=3.0),
                                dict(method="sibson"), dict(method="sibson", sibson_neighbors=50)])
@pytest.mark.parametrize("stream", [2, 1])
def test_stream_kernel_matches_heap_kernel(case_a, tune, kw, stream):
    g, grid, df = case_a
    try:
        set_tuning(**tune)
        _stream_vs_heap(df, grid, stream=stream, **kw)
    finally:
        set_tuning(stream_tile=128, r0=1, ppc=0.5)


@pytest.mark.parametrize("stream", [2, 1])
def test_stream_kernel_sphere_pack_and_fallback_paths(stream):
    n = 48
    mask = synthetic.hex6_sphere_pack_mask(n)
    pts = synthetic.sample_pore_particles(mask, 12000, seed=5)
    vals = synthetic.sphere_pack_flow(pts, n)
    pts, vals, mask = pts.numpy(), vals.numpy(), mask.numpy()
    grid, _ = gi.create_grid(((0, n), (0, n), (0, n)), n)
    # pore voxels only: nearly every tile stays on the streaming kernel
    a, st = _stream_vs_heap(_df(pts, vals), grid, mask=mask, stream=stream, method="idw")
    # nearly every pore voxel is finished by the streaming kernel itself
    # (the statistics describe the last launch: interpolate_field works through the grid in z-chunks)
    assert st["tiles_streamed"] > 0 and 0.8 * int(mask[-48:].sum()) < st["work"]["voxels"] <= int(mask.sum())
    # all voxels: tiles deep inside the grains have no local density estimate -> heap fallback
    b, st2 = _stream_vs_heap(_df(pts, vals), grid, stream=stream, method="idw")
    assert st2["tiles_failed"] > 0
    assert np.abs(a[:, mask] - b[:, mask]).max() <= 1e-11 * np.abs(b).max()
    # lattice + exact duplicates + a pile of coincident points: tie groups overflow the short list
    rng = np.random.default_rng(8)
    lat = np.stack(np.meshgrid(np.arange(10.0), np.arange(10.0), np.arange(10.0), indexing="ij"), -1).reshape(-1, 3)
    cloud = np.concatenate([lat, lat[:200], np.full((120, 3), 4.5),
                            rng.uniform(0, 9, size=(500, 3)).astype(np.float32).astype(np.float64)], 0)
    cv = rng.normal(size=(len(cloud), 3))
    g2, _ = gi.create_grid(((0, 10), (0, 10), (0, 10)), 10)
    c, st3 = _stream_vs_heap(_df(cloud, cv), g2, stream=stream, method="idw", idw_neighbors=30)
    assert st3["tiles_failed"] > 0
    og, _ = rp.create_grid(((0, 10), (0, 10), (0, 10)), 10)
    d, i, _ = rp.knn_bruteforce(cloud, rp.flat_coords(og), 30)
    ref = np.moveaxis(rp.idw_from_knn(d, i, cv).reshape(og[0].shape + (3,)), -1, 0)
    _assert_vel(c, ref, cv)


@pytest.mark.parametrize("k", [64, 130, 300])
def test_large_k(k):
    rng = np.random.default_rng(k)
    pts = rng.uniform(0, 10, size=(700, 3)).astype(np.float32).astype(np.float64)
    vals = rng.normal(size=(700, 3))
    b = ((0, 10), (0, 10), (0, 10))
    grid, _ = gi.create_grid(b, (7, 6, 5))
    U, V, W, kd, ki = gi.interpolate_field(_df(pts, vals), grid, method="idw", idw_neighbors=k, return_knn=True)
    og, _ = rp.create_grid(b, (7, 6, 5))
    Ur, Vr, Wr, d, i = rp.interpolate_field(pts, vals, og, method="idw", idw_neighbors=k, return_knn=True)
    assert np.array_equal(ki, i) and np.array_equal(kd, d)
    _assert_vel(np.stack([U, V, W]), np.stack([Ur, Vr, Wr]), vals)


def test_duplicates_clusters_and_far_queries():
    rng = np.random.default_rng(77)
    base = rng.uniform(0, 4, size=(150, 3)).astype(np.float32).astype(np.float64)
    pts = np.concatenate([base, base[:60], base[:20], np.full((90, 3), 2.0)], 0)  # exact duplicates + a pile
    vals = rng.normal(size=(len(pts), 3))
    b = ((-20, 31), (-3, 9), (1, 5))  # grid reaching far outside the particle bounding box
    res = (17, 6, 4)
    grid, _ = gi.create_grid(b, res)
    U, V, W, kd, ki = gi.interpolate_field(_df(pts, vals), grid, method="idw", idw_neighbors=25, return_knn=True,
                                           out_dtype=np.float64)
    og, _ = rp.create_grid(b, res)
    d, i, _ = rp.knn_bruteforce(pts, rp.flat_coords(og), 25)
    assert np.array_equal(ki, i) and np.array_equal(kd, d)
    ref = rp.idw_from_knn(d, i, vals).reshape(og[0].shape + (3,))
    _assert_vel(np.stack([U, V, W]), np.moveaxis(ref, -1, 0), vals)


def test_known_answers_and_errors():
    rng = np.random.default_rng(5)
    pts = rng.uniform(0, 9, size=(200, 3)).astype(np.float32).astype(np.float64)
    pts[0] = (4.0, 4.0, 4.0)
    grid, _ = gi.create_grid(((0, 10), (0, 10), (0, 10)), 10)
    U, V, W = gi.interpolate_field(_df(pts, np.full((200, 3), 7.0)), grid, method="idw", idw_neighbors=10)
    assert np.allclose(U, 7.0, rtol=1e-6)
    vals = rng.normal(size=(200, 3))
    U, V, W = gi.interpolate_field(_df(pts, vals), grid, method="idw", idw_neighbors=10, out_dtype=np.float64)
    assert abs(U[4, 4, 4] - vals[0, 0]) <= 1e-7 * max(1.0, abs(vals[0, 0]))
    with pytest.raises(IndexError):
        gi.interpolate_field(_df(pts[:5], vals[:5]), grid, method="idw", idw_neighbors=10)
    bad = pts.copy()
    bad[3, 1] = np.nan
    with pytest.raises(ValueError):
        gi.interpolate_field(_df(bad, vals), grid, method="idw", idw_neighbors=10)
    with pytest.raises(ValueError):  # griddata knows 'cubic' only in 1-D / 2-D (interpolator.py:197)
        gi.interpolate_field(_df(pts, vals), grid, method="cubic")
    # dense (np.meshgrid) grids as the reference builds them are accepted too
    og, _ = rp.create_grid(((0, 10), (0, 10), (0, 10)), 10)
    U2, _, _ = gi.interpolate_field(_df(pts, vals), og, method="idw", idw_neighbors=10, out_dtype=np.float64)
    assert np.array_equal(U, U2)


def test_host_cabi_entry_point():
    import ctypes as C
    from ptv_interpolation_b200 import _cabi
    lib = _cabi.load()
    rng = np.random.default_rng(3)
    pts = np.ascontiguousarray(rng.uniform(0, 8, size=(300, 3)).astype(np.float32).astype(np.float64))
    vals = np.ascontiguousarray(rng.normal(size=(300, 3)))
    ax = [np.linspace(0, 7, 8), np.linspace(0, 6, 7), np.linspace(0, 5, 6)]
    out = np.zeros((3, 6, 7, 8), dtype=np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.ptv_interpolate_host(p(pts), p(vals), 300, p(ax[0]), 8, p(ax[1]), 7, p(ax[2]), 6, None,
                                  _cabi.METHOD_IDW, 12, 2.0, 0.0, _cabi.F32, p(out[0]), p(out[1]), p(out[2]))
    _cabi.check(rc)
    Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    Ur, Vr, Wr = rp.interpolate_field(pts, vals, (X, Y, Z), method="idw", idw_neighbors=12)
    _assert_vel(out, np.stack([Ur, Vr, Wr]), vals)
    rc = lib.ptv_interpolate_host(p(pts), p(vals), 5, p(ax[0]), 8, p(ax[1]), 7, p(ax[2]), 6, None,
                                  _cabi.METHOD_IDW, 12, 2.0, 0.0, _cabi.F32, p(out[0]), p(out[1]), p(out[2]))
    assert rc == _cabi.PTV_ERR_TOO_FEW and b"out of bounds" in lib.ptv_last_error()
    # method='linear' (main.py's default) through the same host entry point, float64 output
    out64 = np.zeros((3, 6, 7, 8), dtype=np.float64)
    rc = lib.ptv_interpolate_host(p(pts), p(vals), 300, p(ax[0]), 8, p(ax[1]), 7, p(ax[2]), 6, None,
                                  _cabi.METHOD_LINEAR, 0, 2.0, 0.0, _cabi.F64, p(out64[0]), p(out64[1]), p(out64[2]))
    _cabi.check(rc)
    Ul, Vl, Wl = rp.interpolate_field(pts, vals, (X, Y, Z), method="linear")
    assert np.abs(out64 - np.stack([Ul, Vl, Wl])).max() <= 1e-11
    rc = lib.ptv_interpolate_host(p(pts), p(vals), 4, p(ax[0]), 8, p(ax[1]), 7, p(ax[2]), 6, None,
                                  _cabi.METHOD_LINEAR, 0, 2.0, 0.0, _cabi.F64, p(out64[0]), p(out64[1]), p(out64[2]))
    assert rc == _cabi.PTV_ERR_QHULL and b"QH6214" in lib.ptv_last_error()


# ------------------------------------------------------------------ local RBF (a8/a9)
@pytest.mark.parametrize("name,kw", [("rbf_k20", dict(method="rbf")),
                                     ("rbf_k12_s01", dict(method="rbf", rbf_neighbors=12, smoothing=0.1)),
                                     ("rbf_k40", dict(method="rbf", rbf_neighbors=40)),
                                     ("rbf_cubic_k20", dict(method="rbf", rbf_kernel="cubic")),
                                     ("rbf_linear_k15_s005", dict(method="rbf", rbf_kernel="linear", rbf_neighbors=15,
                                                                  smoothing=0.05)),
                                     ("rbf_quintic_k30", dict(method="rbf", rbf_kernel="quintic", rbf_neighbors=30)),
                                     ("rbf_k60_s001", dict(method="rbf", rbf_neighbors=60, smoothing=0.01))])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_golden_rbf(case_a, name, kw, dtype):
    g, grid, df = case_a
    U, V, W = gi.interpolate_field(df, grid, out_dtype=dtype, **kw)
    _assert_vel(np.stack([U, V, W]), g[name], g["values"])
    if dtype == np.float64:
        err = np.abs(np.stack([U, V, W]) - g[name]).max()
        print(f"rbf {name}: max abs err vs scipy dsysv {err:.3e}")


def test_rbf_reference_own_test(golden_dir):
    """test_parallel.py:6-27: 5 points, rbf with n_jobs=2 (accepted, ignored); neighbours clamp to Np."""
    g = np.load(os.path.join(golden_dir, "case_e_test_parallel.npz"))
    df = pd.DataFrame({"x": [0, 10, 0, 10, 5], "y": [0, 0, 10, 10, 5], "z": [0, 0, 0, 0, 5],
                       "u": [1, 1, 1, 1, 2], "v": [0, 0, 0, 0, 0], "w": [0, 0, 0, 0, 0]})
    grid, _ = gi.create_grid(((0, 10), (0, 10), (0, 10)), 10)
    U, V, W = gi.interpolate_field(df, grid, method="rbf", n_jobs=2, out_dtype=np.float64)
    assert U.shape == (10, 10, 10)
    vals = df[["u", "v", "w"]].values.astype(float)
    vals[:, 1:] = 1.0  # rms scale for the all-zero components
    _assert_vel(np.stack([U, V, W]), g["uvw"], vals)


@pytest.mark.parametrize("k,kernel", [(20, "thin_plate_spline"), (26, "thin_plate_spline"), (28, "cubic"), (14, "quintic"),
                                      (22, "quintic")])
def test_rbf_register_solver_vs_oracle_and_shared_memory_solver(k, kernel):
    """k + tail <= 24 and <= 32 take the two register-resident kernels (rotating Gauss-Jordan); both against SciPy
    (interpolator.py:157-195 -> RBFInterpolator(neighbors=k)) and against the shared-memory elimination."""
    rng = np.random.default_rng(100 + k)
    pts = rng.uniform(0.0, 16.0, size=(900, 3))
    vals = np.stack([np.sin(pts[:, 0] / 3.0), np.cos(pts[:, 1] / 4.0) * pts[:, 2] / 16.0, pts[:, 0] * pts[:, 1] / 200.0], axis=1)
    b = ((1, 15), (1, 15), (1, 15))
    grid, _ = gi.create_grid(b, (9, 8, 7))
    kw = dict(method="rbf", rbf_neighbors=k, rbf_kernel=kernel, out_dtype=np.float64)
    res = np.stack(gi.interpolate_field(_df(pts, vals), grid, **kw))
    og, _ = rp.create_grid(b, (9, 8, 7))
    ref = np.stack(rp.interpolate_field(pts, vals, og, method="rbf", rbf_neighbors=k, rbf_kernel=kernel))
    scale = np.abs(ref).max()
    assert np.abs(res - ref).max() <= 1e-9 * scale
    set_tuning(rbf_regs=0)
    try:
        old = np.stack(gi.interpolate_field(_df(pts, vals), grid, **kw))
    finally:
        set_tuning(rbf_regs=1)
    assert np.abs(res - old).max() <= 1e-9 * scale


def test_rbf_sphere_pack_vs_oracle_and_errors():
    n = 24
    mask = synthetic.hex6_sphere_pack_mask(n)
    pts = synthetic.sample_pore_particles(mask, 2500, seed=21)
    vals = synthetic.sphere_pack_flow(pts, n)
    pts, vals, mask = pts.numpy(), vals.numpy(), mask.numpy()
    b = ((0, n), (0, n), (0, n))
    grid, _ = gi.create_grid(b, n)
    U, V, W = gi.interpolate_field(_df(pts, vals), grid, method="rbf", mask=mask, out_dtype=np.float64)
    og, _ = rp.create_grid(b, n)
    Ur, Vr, Wr = rp.interpolate_field(pts, vals, og, method="rbf")
    ref = np.stack(rp.apply_mask_zero(Ur, Vr, Wr, mask))
    _assert_vel(np.stack([U, V, W]), ref, vals)
    # coplanar neighbourhood: exactly singular saddle-point system -> LinAlgError like scipy
    flat = pts[:300].copy()
    flat[:, 2] = 3.0
    with pytest.raises(np.linalg.LinAlgError):
        gi.interpolate_field(_df(flat, vals[:300]), grid, method="rbf")
    with pytest.raises(ValueError):  # fewer than 4 points: RBFInterpolator raises ValueError
        gi.interpolate_field(_df(pts[:3], vals[:3]), grid, method="rbf")
    with pytest.raises(ValueError):  # documented limit of the CUDA path
        gi.interpolate_field(_df(pts, vals), grid, method="rbf", rbf_neighbors=61)
    with pytest.raises(ValueError, match="epsilon"):  # RBFInterpolator: gaussian needs epsilon, the reference passes none
        gi.interpolate_field(_df(pts, vals), grid, method="rbf", rbf_kernel="gaussian")
    with pytest.raises(ValueError, match="must be one of"):
        gi.interpolate_field(_df(pts, vals), grid, method="rbf", rbf_kernel="bogus")
    with pytest.raises(ValueError, match="At least 10 data points"):
        gi.interpolate_field(_df(pts[:9], vals[:9]), grid, method="rbf", rbf_kernel="quintic")


# ------------------------------------------------------------------ BASELINE-size checks (c2 / c3 geometry)
def _full_size_case(name, device):
    cfg = synthetic.make_config(name, device=device)
    n = cfg["n"]
    ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=device)
    return cfg, n, ax


def test_c2_full_size_planes_vs_oracle_and_properties():
    """Config 2 (256^3, 1M vectors, IDW k=50): three whole z-planes against the oracle (cKDTree over all
    1M particles), plus size-independent properties on the whole grid: constant field -> constant, solid
    voxels exactly zero, slab-by-slab launch == whole-grid launch, divergence/flux consistency."""
    dev = torch.device("cuda", 0)
    cfg, n, ax = _full_size_case("c2", dev)
    eng = PTVEngine(dev)
    mask = cfg["mask"].view(torch.uint8)
    eng.build(cfg["points"], cfg["values"])
    out = eng.interpolate(ax, ax, ax, mask=mask, method="idw", k=50, out_dtype=torch.float64)
    pts, vals = cfg["points"].cpu().numpy(), cfg["values"].cpu().numpy()
    mask_np = cfg["mask"].cpu().numpy()
    assert torch.all(out[:, ~cfg["mask"]] == 0)
    from scipy.spatial import KDTree
    tree = KDTree(pts)
    axn = np.linspace(0, n - 1, n)
    for z in (0, 97, 255):
        Z, Y, X = np.meshgrid(axn[z:z + 1], axn, axn, indexing="ij")
        fc = np.stack([X.ravel(), Y.ravel(), Z.ravel()], -1)
        d, i, _ = rp.knn_canonical(pts, fc, 50, workers=-1, tree=tree)
        ref = np.moveaxis(rp.idw_from_knn(d, i, vals).reshape(1, n, n, 3), -1, 0)
        ref = ref * mask_np[z][None, None]
        _assert_vel(out[:, z:z + 1].cpu().numpy(), ref, vals)
    # slab-by-slab == whole grid (what the z-slab sharding relies on)
    part = eng.interpolate(ax, ax, ax[100:140], mask=mask[100:140], method="idw", k=50, out_dtype=torch.float64)
    assert torch.allclose(part, out[:, 100:140], rtol=0, atol=1e-11)
    # constant field -> constant in the pore space
    eng.build(cfg["points"], torch.full_like(cfg["values"], 3.25))
    const = eng.interpolate(ax, ax, ax, mask=mask, method="idw", k=50)
    pore = const[:, cfg["mask"]]
    assert float((pore - 3.25).abs().max()) <= 1e-5
    # divergence + flux of the float32 field: slab halves with halos == whole grid
    f32 = out.to(torch.float32)
    div, st, qxy, qxz, qyz = eng.divergence_flux(f32[0], f32[1], f32[2], mask, 1.0, 1.0, 1.0)
    h = n // 2
    d0, s0, a0, b0, c0 = eng.divergence_flux(f32[0, :h], f32[1, :h], f32[2, :h], mask[:h], 1.0, 1.0, 1.0,
                                             w_above=f32[2, h].contiguous(), mask_above=mask[h].contiguous())
    d1, s1, a1, b1, c1 = eng.divergence_flux(f32[0, h:], f32[1, h:], f32[2, h:], mask[h:], 1.0, 1.0, 1.0,
                                             w_below=f32[2, h - 1].contiguous())
    assert torch.equal(torch.cat([d0, d1]), div)
    assert torch.allclose(s0 + s1, st, rtol=1e-12) and torch.allclose(torch.cat([a0, a1]), qxy, rtol=1e-12)
    assert torch.allclose(b0 + b1, qxz, rtol=1e-11, atol=1e-9) and torch.allclose(c0 + c1, qyz, rtol=1e-11, atol=1e-9)
    assert torch.allclose(qyz, f32[0].double().sum(dim=(0, 1)), rtol=1e-11, atol=1e-8)


def test_c3_size_sibson_sampled_voxels_vs_oracle():
    """Config 3 geometry (FCC pack, sibson k=50) at 512^3 / 5M vectors: 20k random pore voxels of the
    full-grid result against the oracle, and checksum equality between two runs (determinism)."""
    dev = torch.device("cuda", 0)
    cfg, n, ax = _full_size_case("c3", dev)
    eng = PTVEngine(dev)
    mask = cfg["mask"].view(torch.uint8)
    eng.build(cfg["points"], cfg["values"])
    out = eng.interpolate(ax, ax, ax, mask=mask, method="sibson", k=50, out_dtype=torch.float64)
    chk = out.sum(dtype=torch.float64)
    out2 = eng.interpolate(ax, ax, ax, mask=mask, method="sibson", k=50, out_dtype=torch.float64)
    assert torch.equal(out, out2) and chk == out2.sum(dtype=torch.float64)
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    lin = torch.nonzero(cfg["mask"].reshape(-1)).squeeze(1)
    sel = lin[torch.randint(0, lin.numel(), (20000,), generator=g, device=dev)]
    zz, rem = sel // (n * n), sel % (n * n)
    yy, xx = rem // n, rem % n
    q = torch.stack([xx, yy, zz], -1).to(torch.float64).cpu().numpy()
    pts, vals = cfg["points"].cpu().numpy(), cfg["values"].cpu().numpy()
    d, i, _ = rp.knn_canonical(pts, q, 50, workers=-1)
    ref = rp.sibson_from_knn(d, i, vals).T
    got = out.reshape(3, -1)[:, sel].cpu().numpy()
    _assert_vel(got[:, None, None, :], ref[:, None, None, :], vals)


def test_c4_headline_config_sampled_voxels_vs_oracle():
    """Config 4 -- the benchmarked workload (1024^3 FCC pack, 10M vectors, IDW k=50): 24k random pore voxels
    of the full-grid float32 result against the oracle (cKDTree over all 10M particles, canonical order,
    interpolator.py:126-155), the share of voxels the streaming kernel hands to the exact heap kernel, and
    the z-slab launch the multi-GPU path uses against the whole-grid launch."""
    dev = torch.device("cuda", 0)
    cfg, n, ax = _full_size_case("c4", dev)
    eng = PTVEngine(dev)
    mask = cfg["mask"].view(torch.uint8)
    eng.build(cfg["points"], cfg["values"])
    try:
        set_tuning(stats=1)
        out = eng.interpolate(ax, ax, ax, mask=mask, method="idw", k=50)
        st = eng.knn_stats()
    finally:
        set_tuning(stats=0)
    pore = int(cfg["mask"].sum())
    assert st["used_stream"]
    # the fail list holds 8x4x4 heap tiles: fewer than 0.1 % of the volume is redone by the exact kernel
    assert st["tiles_failed"] * 128 < 1e-3 * n ** 3, st
    assert st["work"]["voxels"] > 0.999 * pore
    assert torch.all(out[:, ~cfg["mask"]] == 0)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    lin = torch.nonzero(cfg["mask"].reshape(-1)).squeeze(1)
    sel = lin[torch.randint(0, lin.numel(), (24000,), generator=g, device=dev)]
    zz, rem = sel // (n * n), sel % (n * n)
    yy, xx = rem // n, rem % n
    q = torch.stack([xx, yy, zz], -1).to(torch.float64).cpu().numpy()
    pts, vals = cfg["points"].cpu().numpy(), cfg["values"].cpu().numpy()
    d, i, _ = rp.knn_canonical(pts, q, 50, workers=-1)
    ref = rp.idw_from_knn(d, i, vals).T
    got = out.reshape(3, -1)[:, sel].cpu().numpy()
    _assert_vel(got[:, None, None, :], ref[:, None, None, :], vals)
    # a z-slab launch (what rank r of N computes) is bit-identical to the same planes of the whole grid
    z0, z1 = 384, 512
    part = eng.interpolate(ax, ax, ax[z0:z1], mask=mask[z0:z1], method="idw", k=50)
    assert torch.equal(part, out[:, z0:z1])


@pytest.mark.parametrize("shape", [(9, 11, 70), (5, 8, 64), (7, 3, 31), (40, 33, 129)])
@pytest.mark.parametrize("thickness", [0, 1, 2, 5])
def test_boundary_particles_bitpacked_vs_oracle(shape, thickness):
    """extract_boundary_particles on the bit-packed path (rows that are not whole words, thickness 0 ==
    binary_dilation until stable) against the oracle port's scipy.ndimage result, element by element."""
    rng = np.random.default_rng(shape[2] + thickness)
    mask = rng.random(shape) > 0.55
    mask[2:4, 1:3, 5:20] = False  # a slab of solid so that thick dilations have something to grow into
    b = ((0, shape[2]), (0, shape[1]), (0, shape[0]))
    for step in (1, 3):
        got = gi.extract_boundary_particles(mask, b, sampling_step=step, thickness=thickness)
        ref = rp.extract_boundary_particles(mask, b, sampling_step=step, thickness=thickness)
        for g_, r_ in zip(got, ref):
            assert np.array_equal(np.asarray(g_), np.asarray(r_))
    empty = gi.extract_boundary_particles(np.zeros(shape, dtype=bool), b, thickness=thickness)
    assert all(len(a) == 0 for a in empty)


def test_sample_mask_on_arbitrary_points_vs_oracle():
    """A grid_tuple that is not a rectilinear meshgrid (the reference accepts any X, Y, Z, interpolator.py:233-236)."""
    rng = np.random.default_rng(5)
    raw = rng.random((12, 10, 14)) > 0.5
    b = ((2.0, 16.0), (0.0, 10.0), (-3.0, 9.0))
    X = rng.uniform(0, 18, size=(4, 5, 6))
    Y = rng.uniform(-2, 12, size=(4, 5, 6))
    Z = rng.uniform(-5, 11, size=(4, 5, 6))
    got = gi.sample_mask_on_grid(raw, (X, Y, Z), b)
    ref = rp.sample_mask_on_grid(raw, (X, Y, Z), b)
    assert got.dtype == np.bool_ and np.array_equal(got, ref)


def test_slab_hash_matches_full_hash_and_detects_short_halos():
    """a rank that bins only its slab's particles plus a halo gets the same neighbours as with
    the whole cloud, and a halo that is too short is detected (the caller then redoes the frame)."""
    dev = torch.device("cuda", 0)
    n = 96
    mask = synthetic.fcc_sphere_pack_mask(n, lattice=32.0, device=dev)
    pts = synthetic.sample_pore_particles(mask, 60000, seed=9)
    vals = synthetic.sphere_pack_flow(pts, n)
    ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)
    m8 = mask.view(torch.uint8)
    eng = PTVEngine(dev)
    for method, k in (("idw", 50), ("sibson", 30), ("idw", 5)):
        eng.build(pts, vals)
        for z0, z1 in ((0, 24), (40, 64), (72, 96)):
            eng.build(pts, vals)
            full, fd, fi = eng.interpolate(ax, ax, ax[z0:z1], mask=m8[z0:z1], method=method, k=k,
                                           out_dtype=torch.float64, return_knn=True)
            assert eng.clip_violations() == 0
            eng.build_slab(pts, vals, float(ax[z0]), float(ax[z1 - 1]), k)
            info = eng.hash_info()
            assert info["dims"][2] < 0.8 * (n / info["cell"])  # fewer cell layers than the whole cloud
            part, pd_, pi = eng.interpolate(ax, ax, ax[z0:z1], mask=m8[z0:z1], method=method, k=k,
                                            out_dtype=torch.float64, return_knn=True)
            assert eng.clip_violations() == 0
            assert torch.equal(pi, fi) and torch.equal(torch.nan_to_num(pd_, nan=-1.0), torch.nan_to_num(fd, nan=-1.0))
            assert torch.allclose(part, full, rtol=0, atol=1e-11)
    # a halo of a fifth of the expected neighbour radius is too short: searches leave the binned range
    eng.build_slab(pts, vals, float(ax[40]), float(ax[63]), 50, halo_factor=0.2)
    eng.interpolate(ax, ax, ax[40:64], mask=m8[40:64], method="idw", k=50)
    assert eng.clip_violations() > 0
    with pytest.raises(ValueError, match="linear"):
        eng.interpolate(ax, ax, ax[40:64], mask=m8[40:64], method="linear")
    eng.build(pts, vals)  # a full build clears the clipping
    eng.interpolate(ax, ax, ax[40:64], mask=m8[40:64], method="idw", k=50)
    assert eng.clip_violations() == 0


def test_production_kernel_neighbour_rows_bitexact_c1_scale():
    """Neighbour rows and distances of the PRODUCTION streaming kernel (not the heap kernel) at config-1
    density: 64^3 hex pack, 12.5k vectors, k = 50 and sibson k = 30, against the canonical cKDTree lists."""
    from ptv_interpolation_b200.engine import default_engine
    n = 64
    mask = synthetic.hex6_sphere_pack_mask(n)
    pts = synthetic.sample_pore_particles(mask, 12500, seed=21)
    vals = synthetic.sphere_pack_flow(pts, n)
    pts, vals, mask = pts.numpy(), vals.numpy(), mask.numpy()
    b = ((0, n), (0, n), (0, n))
    grid, _ = gi.create_grid(b, n)
    og, _ = rp.create_grid(b, n)
    sel = mask.ravel()
    for kw in (dict(method="idw", idw_neighbors=50), dict(method="sibson", sibson_neighbors=30)):
        try:
            set_tuning(stats=1)
            U, V, W, kd, ki = gi.interpolate_field(_df(pts, vals), grid, mask=mask, return_knn=True, **kw)
            st = default_engine().knn_stats()
        finally:
            set_tuning(stats=0)
        # (statistics of the last z-chunk interpolate_field launches: 64 planes -> two chunks of 32)
        assert st["used_stream"] and st["work"]["voxels"] > 0.9 * int(mask[32:].sum())
        Ur, Vr, Wr, d, i = rp.interpolate_field(pts, vals, og, return_knn=True, **kw)
        assert np.array_equal(ki[sel], i[sel]) and np.array_equal(kd[sel], d[sel])
        _assert_vel(np.stack([U, V, W]), np.stack(rp.apply_mask_zero(Ur, Vr, Wr, mask)), vals)


# ------------------------------------------------------------------ N1: outlier filter + point queries
@pytest.mark.parametrize("k,thr", [(25, 3.0), (10, 2.0), (24, 3.5)])
def test_outlier_filter_golden(golden_dir, k, thr):
    from ptv_interpolation_b200 import filtering as gf
    g = np.load(os.path.join(golden_dir, "case_f_filter.npz"))
    kept = gf.remove_outliers_knn(_df(g["points"], g["values"]), k=k, threshold=thr)
    assert np.array_equal(kept[["x", "y", "z", "u", "v", "w"]].values, g[f"kept_k{k}_t{thr}"])
    # kth-neighbour distances (the "filtering radius" report) and the too-small-frame path
    eng = PTVEngine()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    eng.build(t(g["points"]), t(g["values"]))
    keep, kth = eng.outlier_filter(k=k, threshold=thr)
    rk, rkth = rp.outlier_keep_mask(g["points"], g["values"], k, thr)
    assert np.array_equal(keep.cpu().numpy().astype(bool), rk) and np.array_equal(kth.cpu().numpy(), rkth)
    small = _df(g["points"][:5], g["values"][:5])
    assert gf.remove_outliers_knn(small, k=k) is small


def test_apply_filters_namespace():
    import types
    from ptv_interpolation_b200 import filtering as gf
    rng = np.random.default_rng(2)
    pts = rng.uniform(0, 10, size=(500, 3))
    vals = rng.normal(size=(500, 3))
    vals[7] = 50.0
    df = _df(pts, vals)
    args = types.SimpleNamespace(filter_outliers=False, filter_max_speed=10.0, filter_neighbors=20, filter_threshold=3.0)
    assert gf.apply_filters(df, args) is df
    args.filter_outliers = True
    out = gf.apply_filters(df, args)
    thr = df[np.sqrt((df[["u", "v", "w"]].values ** 2).sum(1)) <= 10.0].reset_index(drop=True)
    rk, _ = rp.outlier_keep_mask(thr[["x", "y", "z"]].values, thr[["u", "v", "w"]].values, 20, 3.0)
    assert np.array_equal(out.values, thr[rk].values)


@pytest.mark.parametrize("method,k", [("idw", 12), ("sibson", 30), ("rbf", 20), ("nearest", 1)])
def test_scattered_query_points(method, k):
    """A grid_tuple that is NOT a rectilinear meshgrid (jittered coordinates): the reference treats
    every grid as a point list (interpolator.py:93); so does the point-query kernel."""
    rng = np.random.default_rng(31)
    pts = rng.uniform(0, 12, size=(2500, 3)).astype(np.float32).astype(np.float64)
    vals = rng.normal(size=(2500, 3))
    og, _ = rp.create_grid(((0, 12), (0, 12), (0, 12)), (9, 8, 7))
    grid = tuple(a + 0.3 * rng.random(a.shape) for a in og)
    kw = dict(method=method, idw_neighbors=k, sibson_neighbors=k, rbf_neighbors=k)
    res = gi.interpolate_field(_df(pts, vals), grid, out_dtype=np.float64, return_knn=(method != "rbf"), **kw)
    ref = rp.interpolate_field(pts, vals, grid, return_knn=(method in ("idw", "sibson")), **kw)
    _assert_vel(np.stack(res[:3]), np.stack(ref[:3]), vals)
    if method in ("idw", "sibson"):
        assert np.array_equal(res[4], ref[4]) and np.array_equal(res[3], ref[3])
    # with a mask only the pore points are queried
    m = rng.random(grid[0].shape) > 0.5
    U, V, W = gi.interpolate_field(_df(pts, vals), grid, mask=m, out_dtype=np.float64, **kw)
    full = np.stack(res[:3])
    assert np.all(np.stack([U, V, W])[:, ~m] == 0)
    assert np.abs(np.stack([U, V, W])[:, m] - full[:, m]).max() <= 1e-11 * np.abs(full).max()


# ------------------------------------------------------------------ N3: strain rate / vorticity stencils
def test_strain_vorticity_golden(golden_dir):
    from ptv_interpolation_b200 import velocity_analysis as gva
    g = np.load(os.path.join(golden_dir, "case_g_analysis.npz"))
    dx, dy, dz = (float(h) for h in g["h"])
    s = gva.compute_strain_rate(g["u"], g["v"], g["w"], dx, dy, dz, mask=g["mask"])
    assert s.dtype == np.float64 and np.array_equal(s, g["strain"])
    assert np.array_equal(gva.compute_strain_rate(g["u"], g["v"], g["w"], 1.0, 1.0, 1.0), g["strain_nomask"])
    assert np.array_equal(gva.compute_vorticity(g["u"], g["v"], g["w"], dx, dy, dz, mask=g["mask"]), g["vort"])
    assert np.array_equal(gva.compute_viscous_dissipation(s.copy(), 1.3e-3, mask=g["mask"]), g["diss"])
    u32, v32, w32 = (g[c].astype(np.float32) for c in "uvw")
    s32 = gva.compute_strain_rate(u32, v32, w32, dx, dy, dz, mask=g["mask"])
    ref = rp.compute_strain_rate(u32.astype(np.float64), v32.astype(np.float64), w32.astype(np.float64), dx, dy, dz,
                                 mask=g["mask"])
    assert s32.dtype == np.float32 and np.array_equal(s32, ref.astype(np.float32))
    with pytest.raises(ValueError):
        gva.compute_vorticity(np.zeros((1, 4, 4)), np.zeros((1, 4, 4)), np.zeros((1, 4, 4)), 1, 1, 1)
    # vectorised path (nx % 4 == 0): unit and non-unit spacing, with and without mask, both dtypes
    rng = np.random.default_rng(4)
    for shape, h in (((6, 7, 16), (1.0, 1.0, 1.0)), ((5, 2, 8), (0.7, 2.0, 1.3)), ((2, 9, 4), (2.0, 0.5, 4.0))):
        u, v, w = (rng.normal(size=shape) for _ in range(3))
        m = rng.random(shape) > 0.3
        for mm in (None, m):
            assert np.array_equal(gva.compute_strain_rate(u, v, w, *h, mask=mm), rp.compute_strain_rate(u, v, w, *h, mask=mm))
            assert np.array_equal(gva.compute_vorticity(u, v, w, *h, mask=mm), rp.compute_vorticity(u, v, w, *h, mask=mm))
        uf, vf, wf = (a.astype(np.float32) for a in (u, v, w))
        ref = rp.compute_vorticity(uf.astype(np.float64), vf.astype(np.float64), wf.astype(np.float64), *h, mask=m)
        assert np.array_equal(gva.compute_vorticity(uf, vf, wf, *h, mask=m), ref.astype(np.float32))


@pytest.mark.parametrize("shape,h", [((40, 19, 144), (1.0, 1.0, 1.0)), ((70, 9, 32), (0.5, 2.0, 4.0)),
                                     ((33, 8, 256), (2.00625, 2.00625, 2.00625)), ((3, 3, 16), (1.0, 1.0, 1.0)),
                                     ((66, 17, 128), (1.25, 0.75, 2.0))])
def test_strain_vorticity_bulk_kernel(shape, h):
    """float32 fields with nx % 16 == 0 take the z-marching kernel over bulk-copied plane tiles (strain_bulk.cu):
    several x / y tiles, ragged tiles, z cut into segments, one-sided differences at all six faces; against the
    oracle on the upcast field (velocity_analysis.py:10-63, 94-120) and against the direct-load kernel."""
    from ptv_interpolation_b200 import velocity_analysis as gva
    rng = np.random.default_rng(shape[2] + shape[0])
    u, v, w = (rng.normal(size=shape).astype(np.float32) for _ in range(3))
    m = rng.random(shape) > 0.4
    m[:, :, : shape[2] // 3] &= rng.random() > 0.5  # a solid (or untouched) slab: whole warps without work
    u64, v64, w64 = (a.astype(np.float64) for a in (u, v, w))
    for mm in (m, None):
        s_ref = rp.compute_strain_rate(u64, v64, w64, *h, mask=mm).astype(np.float32)
        o_ref = rp.compute_vorticity(u64, v64, w64, *h, mask=mm).astype(np.float32)
        assert np.array_equal(gva.compute_strain_rate(u, v, w, *h, mask=mm), s_ref)
        assert np.array_equal(gva.compute_vorticity(u, v, w, *h, mask=mm), o_ref)
    eng = PTVEngine()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    s1, o1 = eng.strain_vorticity(t(u), t(v), t(w), *h, mask=t(m))
    set_tuning(stencil_bulk=0)
    try:
        s0, o0 = eng.strain_vorticity(t(u), t(v), t(w), *h, mask=t(m))
    finally:
        set_tuning(stencil_bulk=1)
    assert torch.equal(s0, s1) and torch.equal(o0, o1)


@pytest.mark.parametrize("shape,dtype", [((12, 9, 32), np.float32), ((7, 5, 16), np.float32), ((9, 6, 10), np.float32),
                                         ((8, 4, 16), np.float64)])
def test_strain_vorticity_slabs_with_halos_match_whole(shape, dtype):
    """A grid cut into z-slabs, each with the (3, ny, nx) planes of its z-neighbours: the slab results concatenate to
    the whole-grid result bit for bit (kernel halos for float32 / nx % 16 == 0, padded slabs otherwise)."""
    rng = np.random.default_rng(shape[0] * shape[2])
    u, v, w = (rng.normal(size=shape).astype(dtype) for _ in range(3))
    m = rng.random(shape) > 0.35
    h = (1.25, 0.75, 2.0)
    eng = PTVEngine()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    s_all, o_all = eng.strain_vorticity(t(u), t(v), t(w), *h, mask=t(m))
    cuts = [0, 1, 4, shape[0] - 2, shape[0]]
    s_parts, o_parts = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        below = t(np.stack([u[a - 1], v[a - 1], w[a - 1]])) if a > 0 else None
        above = t(np.stack([u[b], v[b], w[b]])) if b < shape[0] else None
        s, o = eng.strain_vorticity(t(u[a:b]), t(v[a:b]), t(w[a:b]), *h, mask=t(m[a:b]), below=below, above=above)
        s_parts.append(s)
        o_parts.append(o)
    assert torch.equal(torch.cat(s_parts), s_all) and torch.equal(torch.cat(o_parts), o_all)
    ref = rp.compute_strain_rate(u.astype(np.float64), v.astype(np.float64), w.astype(np.float64), *h, mask=m)
    assert np.array_equal(s_all.cpu().numpy(), ref.astype(dtype))


# ------------------------------------------------------------------ N2: projection cleaning
def test_projection_cleaning_golden(golden_dir, capsys):
    g = np.load(os.path.join(golden_dir, "case_h_projection.npz"))
    dx, dy, dz = (float(h) for h in g["h"])
    mask = g["mask"]
    eng = PTVEngine()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    # the matrix-free Laplacian against the reference's CSR matrix, through one LSQR start-up: A^T b
    # is what the first v vector is, so compare via the correction of a known potential instead
    phi_grid = np.zeros(mask.shape)
    phi_grid[mask] = g["lap_x"]
    uc, vc, wc = eng.projection_correct(t(g["u"]), t(g["v"]), t(g["w"]), t(phi_grid), t(mask), dx, dy, dz)
    ru, rv, rw = rp.apply_consistent_correction(g["u"], g["v"], g["w"], g["lap_x"], mask, dx, dy, dz)
    for a, b in ((uc, ru), (vc, rv), (wc, rw)):
        assert np.abs(a.cpu().numpy() - b).max() <= 1e-13
    # one LSQR solve: same stopping reason and iteration count as SciPy, solution within 1e-8
    from scipy.sparse.linalg import lsqr
    A, _ = rp.build_laplacian_matrix(mask, dx, dy, dz)
    b = g["div0"][mask]
    b = b - b.mean()
    ref = lsqr(A, b, damp=1e-8, atol=1e-10, btol=1e-10, iter_lim=3000)
    phi, info = eng.poisson_lsqr(t(g["div0"]), t(mask), dx, dy, dz)
    assert info["istop"] == ref[1] and abs(info["itn"] - ref[2]) <= 2, (info, ref[1:3])
    scale = np.abs(ref[0]).max()
    assert np.abs(phi.cpu().numpy()[mask] - ref[0]).max() <= 1e-7 * scale
    assert np.all(phi.cpu().numpy()[~mask] == 0)
    # anorm accumulates one term per iteration: an iteration count off by one moves it by ~1e-4
    assert abs(info["anorm"] - ref[5]) <= 2e-3 * ref[5] and abs(info["xnorm"] - ref[8]) <= 1e-6 * ref[8]
    # the whole driver (3 iterations) against the reference's output
    u3, v3, w3 = gp.clean_divergence_projection(g["u"], g["v"], g["w"], mask, dx, dy, dz, iterations=3)
    out = capsys.readouterr().out
    assert "DIVERGENCE CLEANING COMPLETE" in out and "Net X-Flux" in out
    for a, b in ((u3, g["u3"]), (v3, g["v3"]), (w3, g["w3"])):
        assert a.dtype == np.float64 and np.abs(a - b).max() <= 1e-7
    u1, v1, w1 = gp.clean_divergence(g["u"], g["v"], g["w"], mask, dx, dy, dz, iterations=1)
    assert np.abs(u1 - g["u1"]).max() <= 1e-7 and np.abs(w1 - g["w1"]).max() <= 1e-7
    with pytest.raises(NotImplementedError):
        gp.clean_divergence(g["u"], g["v"], g["w"], mask, dx, dy, dz, method="variational")


def test_c1_config_vs_c_bruteforce_oracle():
    """Config 1 (hex6 pack 128^3, 100k vectors, IDW k=50) against the SciPy-free C brute-force oracle on
    15k random pore voxels: neighbour lists bit-exact (heap kernel), velocities of the streaming kernel
    within the 1e-5 bar."""
    from oracle.knn_brute import knn_brute
    dev = torch.device("cuda", 0)
    cfg, n, ax = _full_size_case("c1", dev)
    eng = PTVEngine(dev)
    mask = cfg["mask"].view(torch.uint8)
    eng.build(cfg["points"], cfg["values"])
    out = eng.interpolate(ax, ax, ax, mask=mask, method="idw", k=50, out_dtype=torch.float64)
    assert eng.knn_stats()["used_stream"]
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    lin = torch.nonzero(cfg["mask"].reshape(-1)).squeeze(1)
    sel = lin[torch.randperm(lin.numel(), generator=g, device=dev)[:15000]]
    zz, rem = sel // (n * n), sel % (n * n)
    q = torch.stack([rem % n, rem // n, zz], -1).to(torch.float64)
    pts, vals = cfg["points"].cpu().numpy(), cfg["values"].cpu().numpy()
    d, i, _ = knn_brute(pts, q.cpu().numpy(), 50)
    ref = rp.idw_from_knn(d, i, vals).T
    got = out.reshape(3, -1)[:, sel].cpu().numpy()
    _assert_vel(got[:, None, None, :], ref[:, None, None, :], vals)
    # the same voxels as arbitrary query points through the heap kernel: lists bit-exact
    qe = PTVEngine(dev)
    qe.build(q, q)
    _, kd, ki = eng.interpolate_points(qe, method="idw", k=50, return_knn=True)
    assert np.array_equal(ki.cpu().numpy(), i) and np.array_equal(kd.cpu().numpy(), d)


def test_c5_time_resolved_sweep_rebuild_per_frame():
    """Config 5 (frames interpolated back to back, hash rebuilt per frame): one engine reusing its buffers
    over frames of different sizes gives exactly what a fresh engine gives for each frame, and the oracle
    agrees on a sampled frame."""
    dev = torch.device("cuda", 0)
    n = 40
    mask_t = synthetic.fcc_sphere_pack_mask(n, lattice=20.0, device=dev)
    mask = mask_t.view(torch.uint8)
    ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)
    eng = PTVEngine(dev)
    frames = []
    for f, npts in enumerate((9000, 4000, 12000, 9000)):  # growing and shrinking clouds
        pts = synthetic.sample_pore_particles(mask_t, npts, seed=100 + f)
        vals = synthetic.sphere_pack_flow(pts, n) * (1.0 + 0.1 * f)
        eng.build(pts, vals)
        out = eng.interpolate(ax, ax, ax, mask=mask, method="sibson", k=50, out_dtype=torch.float64)
        fresh = PTVEngine(dev)
        fresh.build(pts, vals)
        ref = fresh.interpolate(ax, ax, ax, mask=mask, method="sibson", k=50, out_dtype=torch.float64)
        assert torch.equal(out, ref)
        fresh.close()
        frames.append((pts, vals, out))
    pts, vals, out = frames[2]
    og, _ = rp.create_grid(((0, n), (0, n), (0, n)), n)
    U, V, W = rp.interpolate_field(pts.cpu().numpy(), vals.cpu().numpy(), og, method="sibson", sibson_neighbors=50)
    ref = np.stack(rp.apply_mask_zero(U, V, W, mask_t.cpu().numpy()))
    _assert_vel(out.cpu().numpy(), ref, vals.cpu().numpy())


@pytest.mark.parametrize("res", [(1, 1, 1), (1, 9, 1), (5, 1, 1), (3, 2, 70)])
def test_degenerate_grids_and_clouds(res):
    """Edge cases the reference handles implicitly: one-voxel and line grids, coplanar and collinear
    clouds, k equal to the number of particles, every particle at the same place."""
    rng = np.random.default_rng(sum(res))
    b = ((0, 8), (0, 8), (0, 8))
    grid, _ = gi.create_grid(b, res)
    og, _ = rp.create_grid(b, res)
    clouds = {
        "random": rng.uniform(0, 7, size=(300, 3)),
        "coplanar": np.c_[rng.uniform(0, 7, size=(300, 2)), np.full(300, 2.5)],
        "collinear": np.c_[np.linspace(0, 7, 300), np.full(300, 1.0), np.full(300, 3.0)],
    }
    for name, pts in clouds.items():
        pts = pts.astype(np.float32).astype(np.float64)
        vals = rng.normal(size=(len(pts), 3))
        for method, k in (("idw", 50), ("sibson", 30), ("idw", len(pts)), ("nearest", 1)):
            kw = dict(method=method, idw_neighbors=k, sibson_neighbors=k)
            U, V, W = gi.interpolate_field(_df(pts, vals), grid, out_dtype=np.float64, **kw)
            Ur, Vr, Wr = rp.interpolate_field(pts, vals, og, **kw)
            # where every sibson weight underflows the reference yields NaN, which main.py:195-199 turns into
            # 0; the kernel writes that 0 directly
            # Far from a collinear cloud the k distances are nearly equal, exp(-d/std) lands in the float64
            # DENORMAL range (arguments around -733) and the reference itself only carries ~1e-5 there (it
            # differs from an extended-precision evaluation by 7e-6): compare that one case at 1e-3.
            tol = 1e-3 if (name == "collinear" and method == "sibson") else TOL
            _assert_vel(np.stack([U, V, W]), np.nan_to_num(np.stack([Ur, Vr, Wr])), vals, tol=tol)
        # method='linear': a cloud with volume interpolates (outside the hull: 0); flat clouds have no
        # tetrahedra and Qhull says so (QH6154) -- same exception type here
        from scipy.spatial import QhullError
        if name == "random":
            U, V, W = gi.interpolate_field(_df(pts, vals), grid, method="linear", out_dtype=np.float64)
            Ur, Vr, Wr = rp.interpolate_field(pts, vals, og, method="linear")
            _assert_vel(np.stack([U, V, W]), np.stack([Ur, Vr, Wr]), vals)
        else:
            with pytest.raises(QhullError):
                rp.interpolate_field(pts, vals, og, method="linear")
            with pytest.raises(QhullError):
                gi.interpolate_field(_df(pts, vals), grid, method="linear")
    same = np.full((60, 3), 3.0)
    vals = rng.normal(size=(60, 3))
    U, V, W = gi.interpolate_field(_df(same, vals), grid, method="idw", idw_neighbors=10, out_dtype=np.float64)
    # ten lowest-index particles, all at the same distance: plain mean of their values
    assert np.allclose(U, vals[:10, 0].mean(), rtol=1e-12) and np.allclose(W, vals[:10, 2].mean(), rtol=1e-12)


def test_main_py_flow_steps_4_to_7():
    """The sequence main.py:140-218 runs -- crop, outlier filter, grid, mask resampling (downscale 2),
    no-slip boundary particles, interpolation, solid zeroing, projection cleaning -- through the drop-in
    modules, against the same sequence through the oracle port."""
    import types
    from ptv_interpolation_b200 import filtering as gf
    rng = np.random.default_rng(77)
    nraw = 32
    zz, yy, xx = np.meshgrid(*(np.arange(nraw, dtype=float),) * 3, indexing="ij")
    mask_raw = np.ones((nraw,) * 3, bool)
    for cx, cy, cz, r in ((9, 10, 11, 6.0), (22, 20, 9, 5.0), (15, 24, 23, 6.5)):
        mask_raw &= ((xx - cx) ** 2 + (yy - cy) ** 2 + (zz - cz) ** 2) > r * r
    bounds = ((0, nraw), (0, nraw), (0, nraw))
    p = rng.uniform(-1, nraw + 1, size=(9000, 3)).astype(np.float32).astype(np.float64)
    inside = mask_raw[np.clip(np.rint(p[:, 2]).astype(int), 0, nraw - 1), np.clip(np.rint(p[:, 1]).astype(int), 0, nraw - 1),
                      np.clip(np.rint(p[:, 0]).astype(int), 0, nraw - 1)]
    p = p[inside]
    v = np.stack([0.2 * np.sin(0.3 * p[:, 1]), 0.1 * np.cos(0.2 * p[:, 0]), 1.0 + 0.05 * p[:, 2]], -1)
    v[rng.choice(len(v), 40, replace=False)] *= 6.0  # outliers for the filter
    df = _df(p, v)
    (xmin, xmax), (ymin, ymax), (zmin, zmax) = bounds
    df = df[(df.x >= xmin) & (df.x < xmax) & (df.y >= ymin) & (df.y < ymax) & (df.z >= zmin) & (df.z < zmax)].reset_index(drop=True)
    args = types.SimpleNamespace(filter_outliers=True, filter_max_speed=10.0, filter_neighbors=25, filter_threshold=3.0)

    def flow(mod_i, mod_p, filt):
        d = filt(df)
        res = int(round(nraw / 2))
        (X, Y, Z), (x, y, z) = mod_i.create_grid(bounds, res)
        mask = mod_i.sample_mask_on_grid(mask_raw, (X, Y, Z), bounds)
        bx, by, bz = mod_i.extract_boundary_particles(mask_raw, bounds, sampling_step=3, thickness=2)
        import pandas as pd
        bdf = pd.DataFrame({"x": bx, "y": by, "z": bz, "u": np.zeros_like(bx), "v": np.zeros_like(by), "w": np.zeros_like(bz)})
        d = pd.concat([d, bdf], ignore_index=True)
        return d, (X, Y, Z), (x, y, z), mask

    d_g, grid_g, ax_g, mask_g = flow(gi, gp, lambda d: gf.apply_filters(d, args))
    keep, _ = rp.outlier_keep_mask(df[["x", "y", "z"]].values, df[["u", "v", "w"]].values, 25, 3.0)

    class _RefI:  # the oracle port behind interpolator.py's names
        create_grid = staticmethod(rp.create_grid)
        sample_mask_on_grid = staticmethod(rp.sample_mask_on_grid)
        extract_boundary_particles = staticmethod(rp.extract_boundary_particles)
    d_r, grid_r, ax_r, mask_r = flow(_RefI, None, lambda d: d[keep].reset_index(drop=True))
    assert np.array_equal(d_g.values, d_r.values) and np.array_equal(mask_g, mask_r)
    U, V, W = gi.interpolate_field(d_g, grid_g, method="idw", idw_neighbors=50, idw_power=2.0, out_dtype=np.float64)
    U[~mask_g] = 0; V[~mask_g] = 0; W[~mask_g] = 0  # main.py:202-207 mutates the returned arrays in place
    Ur, Vr, Wr = rp.interpolate_field(d_r[["x", "y", "z"]].values, d_r[["u", "v", "w"]].values, grid_r, method="idw")
    Ur, Vr, Wr = rp.apply_mask_zero(Ur, Vr, Wr, mask_r)
    _assert_vel(np.stack([U, V, W]), np.stack([Ur, Vr, Wr]), d_r[["u", "v", "w"]].values)
    # the masked call gives the same field without the caller's zeroing
    Um, Vm, Wm = gi.interpolate_field(d_g, grid_g, method="idw", mask=mask_g, out_dtype=np.float64)
    assert np.abs(Um - U).max() <= 1e-11 and np.abs(Wm - W).max() <= 1e-11
    # main.py's DEFAULT method (--method linear, main.py:33): same flow, Delaunay interpolation
    Ul, Vl, Wl = gi.interpolate_field(d_g, grid_g, out_dtype=np.float64)
    Ul[~mask_g] = 0; Vl[~mask_g] = 0; Wl[~mask_g] = 0
    Ulr, Vlr, Wlr = rp.interpolate_field(d_r[["x", "y", "z"]].values, d_r[["u", "v", "w"]].values, grid_r, method="linear")
    Ulr, Vlr, Wlr = rp.apply_mask_zero(Ulr, Vlr, Wlr, mask_r)
    _assert_vel(np.stack([Ul, Vl, Wl]), np.stack([Ulr, Vlr, Wlr]), d_r[["u", "v", "w"]].values)
    dx, dy, dz = (float(a[1] - a[0]) for a in ax_g)
    uc, vc, wc = gp.clean_divergence(U, V, W, mask_g, dx, dy, dz, iterations=2)
    ucr, vcr, wcr = rp.clean_divergence_projection(Ur, Vr, Wr, mask_r, dx, dy, dz, iterations=2)
    for a, b in ((uc, ucr), (vc, vcr), (wc, wcr)):
        assert np.abs(a - b).max() <= 1e-6


# ------------------------------------------------------------------ N4: method='linear' (Delaunay, interpolator.py:197)
def _linear_case(g, tag):
    pts, vals = g[tag + "_points"], g[tag + "_values"]
    grid, _ = gi.create_grid(_bounds(g[tag + "_bounds"]), tuple(int(r) for r in g[tag + "_res"]))
    return pts, vals, grid


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("hull", [2, 1, 0])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_linear_golden(golden_dir, tag, hull, dtype):
    """The reference's default method against golden vectors of the unmodified reference: values within
    tolerance, and the tetrahedron each voxel was evaluated in is the one scipy's Delaunay finds (vertex
    rows bit-exact; -1 = outside the convex hull, where the reference's fill_value gives 0)."""
    g = np.load(os.path.join(golden_dir, "case_i_linear.npz"))
    pts, vals, grid = _linear_case(g, tag)
    set_tuning(hull=hull, stats=1)
    try:
        U, V, W, bw, rows = gi.interpolate_field(_df(pts, vals), grid, method="linear", out_dtype=dtype,
                                                 return_knn=True)
        st = gi.default_engine().linear_stats()
    finally:
        set_tuning(hull=2, stats=0)
    assert U.dtype == dtype
    assert st["unresolved"] == 0
    assert np.array_equal(rows, g[tag + "_simplex"])
    _assert_vel(np.stack([U, V, W]), g[tag + "_uvw"], vals)
    outside = g[tag + "_simplex"][:, 0] < 0
    assert np.all(np.stack([U, V, W]).reshape(3, -1)[:, outside] == 0)
    assert st["outside_hull"] == int(outside.sum())
    if dtype == np.float64:
        assert np.abs(np.stack([U, V, W]) - g[tag + "_uvw"]).max() <= 1e-11
        inside = ~outside
        assert np.abs(bw[inside].sum(1) - 1.0).max() <= 1e-12 and bw[inside].min() >= -1e-9


def test_linear_lattice_wall_particles_golden(golden_dir):
    """Pore particles + zero-velocity wall particles on the voxel lattice (main.py:173-178), grid points ON
    lattice sites: co-spherical points make the triangulation non-unique, the interpolant is not (all
    ambiguous tetrahedra carry zeros) -- values match the reference's Qhull result."""
    g = np.load(os.path.join(golden_dir, "case_i_linear.npz"))
    cb = np.load(os.path.join(golden_dir, "case_b_boundary.npz"))
    grid, _ = gi.create_grid(((0, 12), (0, 12), (0, 12)), 12)
    set_tuning(stats=1)
    try:
        U, V, W = gi.interpolate_field(_df(cb["points"], cb["values"]), grid, method="linear", out_dtype=np.float64)
        st = gi.default_engine().linear_stats()
    finally:
        set_tuning(stats=0)
    assert st["unresolved"] == 0
    assert np.abs(np.stack([U, V, W]) - g["c_uvw"]).max() <= 1e-9
    # with the pore mask, solid voxels are skipped and written 0 (what main.py:202-207 does next)
    m = cb["mask_grid"]
    Um, Vm, Wm = gi.interpolate_field(_df(cb["points"], cb["values"]), grid, method="linear", mask=m,
                                      out_dtype=np.float64)
    assert np.all(np.stack([Um, Vm, Wm])[:, ~m] == 0)
    # (a different warp bounding box may pick another of the equivalent tetrahedra: same value to rounding)
    assert np.abs(np.stack([Um, Vm, Wm])[:, m] - np.stack([U, V, W])[:, m]).max() <= 1e-11


@pytest.mark.parametrize("hull", [2, 1, 0])
def test_linear_sphere_pack_vs_oracle(hull):
    """Config-1-like sphere pack (96^3 grid, 40k vectors + lattice wall particles, pore mask): every pore
    voxel against scipy's Delaunay (simplex rows exact, values within tolerance); linear precision: a
    field that is linear in space is reproduced exactly inside the hull."""
    n = 72
    mask = synthetic.hex6_sphere_pack_mask(n)
    p = synthetic.sample_pore_particles(mask, 30000, seed=5)
    v = synthetic.sphere_pack_flow(p, n)
    p, v, mask = p.numpy(), v.numpy(), mask.numpy()
    bounds = ((0, n), (0, n), (0, n))
    bx, by, bz = rp.extract_boundary_particles(mask, bounds, sampling_step=3, thickness=1)
    pts = np.concatenate([p, np.stack([bx, by, bz], -1)], 0)
    vals = np.concatenate([v, np.zeros((len(bx), 3))], 0)
    grid, _ = gi.create_grid(bounds, n)
    set_tuning(hull=hull, stats=1)
    try:
        U, V, W, bw, rows = gi.interpolate_field(_df(pts, vals), grid, method="linear", mask=mask,
                                                 out_dtype=np.float64, return_knn=True)
        st = gi.default_engine().linear_stats()
    finally:
        set_tuning(hull=2, stats=0)
    assert st["unresolved"] == 0
    og, _ = rp.create_grid(bounds, n)
    fc = rp.flat_coords(og)
    sel = np.flatnonzero(mask.ravel())
    ref_rows, ref_b = rp.delaunay_simplex_rows(pts, fc[sel])
    same = (rows[sel] == ref_rows).all(1)
    # the lattice wall particles are co-spherical: where Qhull and the kernel split such a cell differently
    # (0.1 % of the voxels) the tetrahedra differ but, all ambiguous vertices carrying zeros, the values do not
    assert same.mean() >= 0.995, float(same.mean())
    ref = rp.interpolate_field(pts, vals, (fc[sel, 0], fc[sel, 1], fc[sel, 2]), method="linear")
    got = np.stack([U, V, W]).reshape(3, -1)[:, sel]
    _assert_vel(got, np.stack(ref), vals)
    assert np.all(np.stack([U, V, W])[:, ~mask] == 0)
    lin = np.stack([2.0 + 0.5 * pts[:, 0], -1.0 + 0.25 * pts[:, 1] - 0.1 * pts[:, 2], 0.3 * pts[:, 2]], -1)
    Ul, Vl, Wl = gi.interpolate_field(_df(pts, lin), grid, method="linear", mask=mask, out_dtype=np.float64)
    inside = ref_rows[:, 0] >= 0
    q = fc[sel][inside]
    gotl = np.stack([Ul, Vl, Wl]).reshape(3, -1)[:, sel][:, inside]
    expect = np.stack([2.0 + 0.5 * q[:, 0], -1.0 + 0.25 * q[:, 1] - 0.1 * q[:, 2], 0.3 * q[:, 2]])
    assert np.abs(gotl - expect).max() <= 1e-9


def test_linear_scattered_points_duplicates_and_errors():
    """Non-rectilinear query points go through the point-query form; too few particles raise what Qhull
    raises; the all-particles candidate path (tiny clouds, dense clouds per voxel) agrees with scipy."""
    from scipy.spatial import QhullError
    rng = np.random.default_rng(77)
    pts = rng.uniform(0, 12, size=(3000, 3)).astype(np.float32).astype(np.float64)
    vals = rng.normal(size=(3000, 3))
    og, _ = rp.create_grid(((0, 12), (0, 12), (0, 12)), (9, 8, 7))
    grid = tuple(a + 0.3 * rng.random(a.shape) for a in og)
    res = gi.interpolate_field(_df(pts, vals), grid, method="linear", out_dtype=np.float64, return_knn=True)
    ref = rp.interpolate_field(pts, vals, grid, method="linear")
    _assert_vel(np.stack(res[:3]), np.stack(ref), vals)
    ref_rows, _ = rp.delaunay_simplex_rows(pts, rp.flat_coords(grid))
    assert np.array_equal(res[4], ref_rows)
    with pytest.raises(QhullError):
        gi.interpolate_field(_df(pts[:4], vals[:4]), og, method="linear")
    # a dense cloud on a coarse grid (hundreds of particles per voxel) and a tiny cloud
    for n, r in ((60000, (6, 5, 4)), (9, (7, 6, 5))):
        p2 = rng.uniform(0, 12, size=(n, 3)).astype(np.float32).astype(np.float64)
        v2 = rng.normal(size=(n, 3))
        g2, _ = gi.create_grid(((0.5, 12.5), (0.5, 12.5), (0.5, 12.5)), r)
        o2, _ = rp.create_grid(((0.5, 12.5), (0.5, 12.5), (0.5, 12.5)), r)
        U, V, W = gi.interpolate_field(_df(p2, v2), g2, method="linear", out_dtype=np.float64)
        _assert_vel(np.stack([U, V, W]), np.stack(rp.interpolate_field(p2, v2, o2, method="linear")), v2)


@pytest.mark.parametrize("case", ["quasi2d", "offset", "clustered"])
def test_linear_anisotropic_offset_and_clustered_clouds(case):
    """Shapes that stress the search rather than the arithmetic: the quasi-2-D slab of generate_cylinders.py
    (size x size/2 x 16), a cloud far from the origin, and a cloud whose density varies by 1000x."""
    rng = np.random.default_rng({"quasi2d": 1, "offset": 2, "clustered": 3}[case])
    if case == "quasi2d":
        pts = rng.uniform([0, 0, 0], [64, 32, 4], size=(20000, 3))
        b, res = ((0, 65), (0, 33), (0, 5)), (40, 20, 5)
    elif case == "offset":
        off = np.array([1.0e4, -2.0e4, 3.0e4])
        pts = rng.uniform(0, 20, size=(8000, 3)) + off
        b = tuple((float(o) + 1.0, float(o) + 20.0) for o in off)
        res = (17, 16, 15)
    else:
        dense = rng.normal([6, 6, 6], 0.4, size=(15000, 3))
        sparse = rng.uniform(0, 24, size=(1500, 3))
        pts = np.concatenate([dense, sparse])
        b, res = ((0, 25), (0, 25), (0, 25)), (24, 23, 22)
    vals = np.stack([np.sin(0.3 * pts[:, 0]), 0.1 * pts[:, 1], np.cos(0.2 * pts[:, 2])], -1) - 0.5
    grid, _ = gi.create_grid(b, res)
    og, _ = rp.create_grid(b, res)
    set_tuning(stats=1)
    try:
        U, V, W, bw, rows = gi.interpolate_field(_df(pts, vals), grid, method="linear", out_dtype=np.float64,
                                                 return_knn=True)
        st = gi.default_engine().linear_stats()
    finally:
        set_tuning(stats=0)
    assert st["unresolved"] == 0
    if case == "offset":
        # Qhull works on absolute coordinates: 3e4 away from the origin its own answer moves (one voxel of
        # this grid changes by 5.5e-4 against the same cloud shifted back -- the SciPy-free restatement
        # oracle/delaunay_lp.py sides with the shifted-back answer to 1e-16).  The kernel works on differences
        # and is translation invariant, so it is compared with the reference on the shifted-back problem.
        pts_r, og_r = pts - off, tuple(a - o for a, o in zip(og, off))
        assert np.array_equal(pts_r + off, pts)
    else:
        pts_r, og_r = pts, og
    ref = np.stack(rp.interpolate_field(pts_r, vals, og_r, method="linear"))
    _assert_vel(np.stack([U, V, W]), ref, vals)
    ref_rows, _ = rp.delaunay_simplex_rows(pts_r, rp.flat_coords(og_r))
    assert (rows == ref_rows).all(1).mean() >= 0.999
    assert np.array_equal(rows[:, 0] < 0, ref_rows[:, 0] < 0)  # the same voxels lie outside the hull
