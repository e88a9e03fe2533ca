"""z-slab sharding of the voxel grid across ranks (one process per GPU, torch.distributed).

The interpolation itself needs no data-path collective: every rank holds the particle cloud and
interpolates its own contiguous block of z-planes.  The exchange steps are the ones the path
really has: one-plane halos of ``w`` and ``mask`` between z-neighbours for the
z-term of the divergence stencil (physics.py:49-53), and sum-reductions of the flux profiles and
of (sum|div|, n_fluid).  Works on NCCL (CUDA tensors) and gloo (CPU tensors, used by the tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["slab_range", "slab_range_weighted", "SlabComm"]


def slab_range(nz: int, world: int, rank: int):
    """Contiguous z-planes [z0, z1) of ``rank``: the first ``nz % world`` ranks get one extra."""
    base, rem = divmod(nz, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def slab_range_weighted(weights, world: int, rank: int):
    """Contiguous z-planes [z0, z1) of ``rank`` such that every rank gets about the same share of
    ``sum(weights)`` (weights[z] = work of plane z, e.g. its pore-voxel count: solid voxels cost nothing) and
    at least one plane.  Deterministic: every rank computes the same cuts from the same weights."""
    w = [float(x) for x in weights]
    nz = len(w)
    if world > nz:
        raise ValueError("more ranks than z-planes")
    total = sum(w)
    if total <= 0.0:
        return slab_range(nz, world, rank)
    cuts, acc, z = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        # advance while the plane's midpoint lies below the target; keep one plane for every rank on both sides
        while z < nz - (world - r) and (z < cuts[-1] + 1 or acc + 0.5 * w[z] < target):
            acc += w[z]
            z += 1
        cuts.append(z)
    cuts.append(nz)
    return cuts[rank], cuts[rank + 1]


class SlabComm:
    """Halo exchange + reductions for one slab decomposition.  With world == 1 every method is a
    no-op that returns the local data."""

    def __init__(self, nz: int, group=None, plane_weights=None):
        """``plane_weights`` (length nz, identical on every rank; e.g. pore voxels per z-plane) balances the
        slabs by work instead of by plane count."""
        self.group = group
        self.on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.nz = nz
        if plane_weights is not None and self.world > 1:
            if len(plane_weights) != nz:
                raise ValueError("plane_weights must have one entry per z-plane")
            self.z0, self.z1 = slab_range_weighted(plane_weights, self.world, self.rank)
        else:
            self.z0, self.z1 = slab_range(nz, self.world, self.rank)
        # neighbours that own at least one plane
        self.lower = self.rank - 1 if self.rank > 0 else None
        self.upper = self.rank + 1 if self.rank + 1 < self.world else None

    def exchange_halos(self, w_slab: torch.Tensor, mask_slab: torch.Tensor):
        """Returns (w_below, w_above, mask_above): the (ny,nx) planes z0-1 and z1 owned by the
        z-neighbours (None at the domain edges -> Neumann edge rule of physics.py:38-45)."""
        return self.wait_halos(self.post_halos(w_slab[0], w_slab[-1], mask_slab[0]))

    def post_halos(self, first_w: torch.Tensor, last_w: torch.Tensor, first_m: torch.Tensor):
        """Start the halo exchange as soon as the slab's first and last planes exist (the planes in between
        may still be being interpolated); wait_halos() returns the neighbours' planes."""
        if self.world == 1:
            return None
        if self.z1 - self.z0 < 1:
            raise ValueError("every rank must own at least one z-plane")
        ops, w_below, w_above, m_above = [], None, None, None
        first_w = first_w.contiguous()
        last_w = last_w.contiguous()
        first_m = first_m.contiguous()
        if self.lower is not None:  # my first plane is the lower neighbour's "above"
            w_below = torch.empty_like(first_w)
            ops += [dist.P2POp(dist.isend, first_w, self.lower, self.group),
                    dist.P2POp(dist.isend, first_m, self.lower, self.group),
                    dist.P2POp(dist.irecv, w_below, self.lower, self.group)]
        if self.upper is not None:
            w_above = torch.empty_like(last_w)
            m_above = torch.empty_like(first_m)
            ops += [dist.P2POp(dist.isend, last_w, self.upper, self.group),
                    dist.P2POp(dist.irecv, w_above, self.upper, self.group),
                    dist.P2POp(dist.irecv, m_above, self.upper, self.group)]
        reqs = dist.batch_isend_irecv(ops)
        return reqs, (w_below, w_above, m_above), (first_w, last_w, first_m)  # the send buffers stay alive

    def wait_halos(self, handle):
        if handle is None:
            return None, None, None
        for req in handle[0]:
            req.wait()
        return handle[1]

    def exchange_planes(self, first: torch.Tensor, last: torch.Tensor):
        """Generic one-plane halo exchange: ``first`` / ``last`` are this rank's outermost planes (any shape, e.g. the
        (3, ny, nx) stack of u, v, w for the gradient stencils).  Returns (below, above): the lower neighbour's
        ``last`` and the upper neighbour's ``first`` (None at the domain faces)."""
        if self.world == 1:
            return None, None
        first, last = first.contiguous(), last.contiguous()
        ops, below, above = [], None, None
        if self.lower is not None:
            below = torch.empty_like(first)
            ops += [dist.P2POp(dist.isend, first, self.lower, self.group), dist.P2POp(dist.irecv, below, self.lower, self.group)]
        if self.upper is not None:
            above = torch.empty_like(last)
            ops += [dist.P2POp(dist.isend, last, self.upper, self.group), dist.P2POp(dist.irecv, above, self.upper, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return below, above

    def reduce_profiles_(self, acc: torch.Tensor):
        """ONE sum-all-reduce for everything the stencil pass accumulates: ``acc`` is the flat float64 buffer
        [sum|div|, n_fluid | Q_xy[nz] (each rank fills its own planes, zeros elsewhere) | Q_xz[ny] | Q_yz[nx]]."""
        if self.world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
        return acc

    def reduce_sum_(self, *tensors):
        """In-place SUM all-reduce of small float64 tensors (flux profiles, statistics)."""
        if self.world > 1:
            flat = torch.cat([t.reshape(-1) for t in tensors])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            o = 0
            for t in tensors:
                n = t.numel()
                t.copy_(flat[o:o + n].view_as(t))
                o += n
        return tensors

    def gather_planes(self, q_local: torch.Tensor):
        """Concatenate a per-plane profile (e.g. Q_xy[z] of the local slab) over ranks."""
        if self.world == 1:
            return q_local
        full = torch.zeros(self.nz, dtype=q_local.dtype, device=q_local.device)
        full[self.z0:self.z1] = q_local
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
        return full
