"""world_size-2 (and 3) gloo tests of the z-slab sharding logic on CPU tensors: slab ranges, halo
exchange, flux/statistics reductions.  The divergence of each slab is evaluated with the oracle's
closed form on the halo-extended slab, which must reproduce the whole-grid result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reference_port as rp
from ptv_interpolation_b200.distributed import SlabComm, slab_range, slab_range_weighted


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_slab_range_partitions():
    for nz in (1, 7, 8, 128, 1000):
        for world in (1, 2, 3, 8):
            if world > nz:
                continue
            cuts = [slab_range(nz, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == nz
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_slab_range_weighted_balances_work():
    rng = np.random.default_rng(0)
    for nz, world in ((16, 2), (100, 8), (1024, 8), (9, 9), (40, 3)):
        w = rng.random(nz) * (1.0 + np.sin(np.arange(nz) / 3.0) ** 2)
        cuts = [slab_range_weighted(w, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == nz
        assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
        assert all(b > a for a, b in cuts)
        if nz >= 10 * world:
            work = np.array([w[a:b].sum() for a, b in cuts])
            assert work.max() <= w.sum() / world + w.max()  # within one plane of the ideal share
    assert slab_range_weighted(np.zeros(12), 3, 1) == slab_range(12, 3, 1)
    with pytest.raises(ValueError):
        slab_range_weighted(np.ones(3), 4, 0)


def _worker(rank, world, port, shape, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # same data on every rank
        u, v, w = (rng.normal(size=shape) for _ in range(3))
        m = rng.random(shape) > 0.3
        nz = shape[0]
        comm = SlabComm(nz)
        z0, z1 = comm.z0, comm.z1
        assert (z0, z1) == slab_range(nz, world, rank)
        wt = torch.from_numpy(w[z0:z1].copy())
        mt = torch.from_numpy(m[z0:z1].astype(np.uint8))
        w_below, w_above, m_above = comm.exchange_halos(wt, mt)
        if rank == 0:
            assert w_below is None
        else:
            assert np.array_equal(w_below.numpy(), w[z0 - 1])
        if rank == world - 1:
            assert w_above is None and m_above is None
        else:
            assert np.array_equal(w_above.numpy(), w[z1]) and np.array_equal(m_above.numpy(), m[z1].astype(np.uint8))
        # slab divergence from the halo-extended slab == whole-grid divergence
        lo = 1 if w_below is not None else 0
        hi = 1 if w_above is not None else 0
        ext = slice(z0 - lo, z1 + hi)
        d = rp.compute_consistent_divergence(u[ext], v[ext], w[ext], m[ext], 1.0, 2.0, 0.5)
        d = d[lo:d.shape[0] - hi]
        ref = rp.compute_consistent_divergence(u, v, w, m, 1.0, 2.0, 0.5)
        assert np.array_equal(d, ref[z0:z1])
        # reductions
        q_xy = torch.from_numpy(w[z0:z1].sum(axis=(1, 2)))
        q_xz = torch.from_numpy(v[z0:z1].sum(axis=(0, 2)))
        q_yz = torch.from_numpy(u[z0:z1].sum(axis=(0, 1)))
        stats = torch.tensor([np.abs(d[m[z0:z1]]).sum(), float(m[z0:z1].sum())], dtype=torch.float64)
        comm.reduce_sum_(q_xz, q_yz, stats)
        full_xy = comm.gather_planes(q_xy)
        assert np.allclose(full_xy.numpy(), w.sum(axis=(1, 2)), rtol=1e-13, atol=1e-13)
        assert np.allclose(q_xz.numpy(), v.sum(axis=(0, 2)), rtol=1e-12, atol=1e-12)
        assert np.allclose(q_yz.numpy(), u.sum(axis=(0, 1)), rtol=1e-12, atol=1e-12)
        assert abs(stats[0].item() / stats[1].item() - rp.mean_abs_div(ref, m)) < 1e-13
        # the pipeline's form: halos posted from the boundary planes alone, ONE all-reduce over the flat
        # accumulator [sum|div|, n_fluid | Q_xy (own planes, zeros elsewhere) | Q_xz | Q_yz]
        pend = comm.post_halos(wt[0], wt[-1], mt[0])
        wb2, wa2, ma2 = comm.wait_halos(pend)
        for a2, b2 in ((wb2, w_below), (wa2, w_above), (ma2, m_above)):
            assert (a2 is None and b2 is None) or torch.equal(a2, b2)
        ny, nx = shape[1], shape[2]
        acc = torch.zeros(2 + nz + ny + nx, dtype=torch.float64)
        acc[0] = float(np.abs(d[m[z0:z1]]).sum())
        acc[1] = float(m[z0:z1].sum())
        acc[2 + z0:2 + z1] = torch.from_numpy(w[z0:z1].sum(axis=(1, 2)))
        acc[2 + nz:2 + nz + ny] = torch.from_numpy(v[z0:z1].sum(axis=(0, 2)))
        acc[2 + nz + ny:] = torch.from_numpy(u[z0:z1].sum(axis=(0, 1)))
        comm.reduce_profiles_(acc)
        assert np.allclose(acc[2:2 + nz].numpy(), w.sum(axis=(1, 2)), rtol=1e-13, atol=1e-13)
        assert np.allclose(acc[2 + nz:2 + nz + ny].numpy(), v.sum(axis=(0, 2)), rtol=1e-12, atol=1e-12)
        assert np.allclose(acc[2 + nz + ny:].numpy(), u.sum(axis=(0, 1)), rtol=1e-12, atol=1e-12)
        assert abs(acc[0].item() / acc[1].item() - rp.mean_abs_div(ref, m)) < 1e-13
        # gradient stencils on slabs (velocity_analysis.py:10-120): one (3, ny, nx) plane of u, v, w from each
        # z-neighbour makes the slab's result equal to the whole-grid result
        stack = torch.from_numpy(np.stack([u[z0:z1], v[z0:z1], w[z0:z1]]))
        below, above = comm.exchange_planes(stack[:, 0], stack[:, -1])
        assert (below is None) == (rank == 0) and (above is None) == (rank == world - 1)
        if below is not None:
            assert np.array_equal(below.numpy(), np.stack([u[z0 - 1], v[z0 - 1], w[z0 - 1]]))
        if above is not None:
            assert np.array_equal(above.numpy(), np.stack([u[z1], v[z1], w[z1]]))
        if z1 - z0 + lo + hi >= 2:
            s_ext = rp.compute_strain_rate(u[ext], v[ext], w[ext], 1.0, 2.0, 0.5)
            s_ref = rp.compute_strain_rate(u, v, w, 1.0, 2.0, 0.5)
            assert np.array_equal(s_ext[lo:s_ext.shape[0] - hi], s_ref[z0:z1])
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (9, 6, 5)), (3, (7, 4, 6))])
def test_slab_comm_gloo(tmp_path, world, shape):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, shape, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_single_process_comm_is_noop():
    comm = SlabComm(10)
    assert (comm.z0, comm.z1, comm.world) == (0, 10, 1)
    assert comm.exchange_halos(torch.zeros(10, 2, 2), torch.zeros(10, 2, 2, dtype=torch.uint8)) == (None, None, None)
    t = torch.ones(3, dtype=torch.float64)
    comm.reduce_sum_(t)
    assert torch.equal(t, torch.ones(3, dtype=torch.float64))
    assert torch.equal(comm.gather_planes(t), t)
