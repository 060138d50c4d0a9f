"""Multi-GPU execution of the hot path on one 8 x B200 box: one process per GPU,
``torch.distributed`` (NCCL over NVLink / NVSwitch) for the plumbing.

Two partitionings:

* **ensemble sharding** -- rows of an ensemble are independent problems (own ``dt``, own
  termination): block-partition the batch axis, no collective in the time loop
  (:func:`shard_rows`, :class:`ShardedEnsemble`);
* **slab decomposition** of one very large periodic grid -- contiguous slabs of ``n / G`` cells
  per GPU on a ring; every RHS needs the 3 cells next to each slab edge from the neighbour
  (24 B per side per stage: latency, not bandwidth), and a CFL-adaptive ``dt`` needs one
  ``all_reduce(MAX)`` of a single double per step (:class:`SlabSolver`).

The reference has no distributed path at all (platform forced to one CPU device,
``pyshocks/__init__.py:66``); these are the new workloads of BASELINE.json configs 3-5.
"""

from __future__ import annotations

from typing import Protocol, Sequence

import torch
import torch.distributed as dist

from .ensemble import EnsembleSolver, SolveResult


def shard_rows(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Block partition of ``batch`` rows: ``(first_row, rows)`` of ``rank`` (remainder rows go to
    the lowest ranks, so shard sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(batch, world)
    rows = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, rows


class ShardedEnsemble:
    """This rank's block of an ensemble; the time loop is :class:`EnsembleSolver`'s, untouched."""

    def __init__(self, *, batch: int, rank: int | None = None, world: int | None = None, **solver_kwargs) -> None:
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if dist.is_initialized() else 0)
        self.global_batch = batch
        self.first, self.rows = shard_rows(batch, self.rank, self.world)
        self.solver = EnsembleSolver(batch=self.rows, **solver_kwargs)

    def local_rows(self, u_global: torch.Tensor) -> torch.Tensor:
        return u_global[self.first : self.first + self.rows]

    def gather(self, dst: int = 0) -> torch.Tensor | None:
        """Collect every rank's final state on ``dst`` (one collective, outside the time loop)."""
        local = self.solver.u.contiguous()
        if self.world == 1:
            return local
        shapes = [shard_rows(self.global_batch, r, self.world)[1] for r in range(self.world)]
        # NCCL collectives want equal shapes: pad every shard to the largest one
        rows_max = max(shapes)
        padded = torch.zeros((rows_max, local.shape[1]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
        bufs = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(bufs, padded)
        if self.rank != dst:
            return None
        return torch.cat([b[:rows] for b, rows in zip(bufs, shapes)], dim=0)


# {{{ halo exchange on a periodic ring


class Ring(Protocol):
    rank: int
    world: int

    def exchange(self, send_left: torch.Tensor, send_right: torch.Tensor,
                 recv_left: torch.Tensor, recv_right: torch.Tensor) -> None: ...

    def all_max(self, x: torch.Tensor) -> None: ...


class DistRing:
    """Neighbour exchange through ``torch.distributed`` point-to-point ops (NCCL send/recv pairs
    grouped in one launch; gloo on CPU for the tests)."""

    def __init__(self, group: dist.ProcessGroup | None = None) -> None:
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange(self, send_left, send_right, recv_left, recv_right) -> None:
        left = (self.rank - 1) % self.world
        right = (self.rank + 1) % self.world
        if self.world == 1:
            recv_left.copy_(send_right)
            recv_right.copy_(send_left)
            return
        ops = [
            dist.P2POp(dist.isend, send_left, left, self.group),
            dist.P2POp(dist.isend, send_right, right, self.group),
            dist.P2POp(dist.irecv, recv_left, left, self.group),
            dist.P2POp(dist.irecv, recv_right, right, self.group),
        ]
        if self.world == 2:
            # left and right neighbour are the same rank: order the pairs by tag so that the
            # two messages cannot be swapped
            ops = [
                dist.P2POp(dist.isend, send_left, left, self.group, tag=1),
                dist.P2POp(dist.irecv, recv_right, right, self.group, tag=1),
                dist.P2POp(dist.isend, send_right, right, self.group, tag=2),
                dist.P2POp(dist.irecv, recv_left, left, self.group, tag=2),
            ]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def all_max(self, x: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(x, op=dist.ReduceOp.MAX, group=self.group)


# }}}


def fill_halos_local(states: Sequence[torch.Tensor], g: int) -> None:
    """Periodic ring halo fill between slab arrays ``[g | n_r | g]`` held in one process."""
    world = len(states)
    for r, u in enumerate(states):
        left, right = states[(r - 1) % world], states[(r + 1) % world]
        u[..., :g] = left[..., left.shape[-1] - 2 * g : left.shape[-1] - g]
        u[..., u.shape[-1] - g :] = right[..., g : 2 * g]


def fill_halos(ring: DistRing, u: torch.Tensor, g: int, scratch: torch.Tensor) -> None:
    """Periodic ring halo fill of this rank's slab array ``[g | n_local | g]`` (last axis).

    ``scratch``: ``(4, rows, g)`` staging buffer (contiguous send / receive blocks)."""
    nx = u.shape[-1]
    send_l, send_r, recv_l, recv_r = scratch[0], scratch[1], scratch[2], scratch[3]
    send_l.copy_(u[..., g : 2 * g].reshape(send_l.shape))            # my first g interior cells -> left neighbour's right halo
    send_r.copy_(u[..., nx - 2 * g : nx - g].reshape(send_r.shape))  # my last g interior cells -> right neighbour's left halo
    ring.exchange(send_l, send_r, recv_l, recv_r)
    u[..., :g] = recv_l.reshape(u[..., :g].shape)
    u[..., nx - g :] = recv_r.reshape(u[..., nx - g :].shape)


class SlabSolver:
    """One rank's slab of a single periodic 1-D Burgers grid of ``n_global`` cells.

    Per stage: ring halo exchange of 3 cells per side, then the fused stage kernel with the
    boundary kind "none" (ghost cells used as found).  ``dt`` is either fixed or CFL-adaptive
    with one ``all_reduce(MAX)`` of the fused per-rank maximum per step."""

    def __init__(self, *, n_global: int, ring: DistRing, flux: str = "rusanov", rec: str = "wenojs53",
                 dx: float, eps: float = 1.0e-12, math: str = "fast", device: torch.device | str | None = None) -> None:
        self.ring = ring
        self.first, self.n_local = shard_rows(n_global, ring.rank, ring.world)
        self.g = {"constant": 1, "wenojs32": 2, "wenojs53": 3}[rec]
        if self.n_local < 2 * self.g:
            raise ValueError("slabs must hold at least 2 g cells")
        self.solver = EnsembleSolver(equation="burgers", flux=flux, rec=rec, bc="none", n=self.n_local, g=self.g,
                                     dx=dx, eps=eps, batch=1, math=math, device=device)
        dev = self.solver.hp.device
        self.scratch = torch.zeros((4, 1, self.g), dtype=torch.float64, device=dev)
        self.exchanges = 0

    @property
    def u(self) -> torch.Tensor:
        return self.solver.u

    def load_interior(self, u_local: torch.Tensor) -> None:
        """``u_local``: this rank's ``n_local`` interior cells."""
        self.solver.u[0, self.g : self.g + self.n_local].copy_(u_local)

    def interior(self) -> torch.Tensor:
        return self.solver.u[0, self.g : self.g + self.n_local]

    def _halo(self, u: torch.Tensor) -> None:
        fill_halos(self.ring, u, self.g, self.scratch)
        self.exchanges += 1

    def step(self, dt: torch.Tensor, maxabs: torch.Tensor | None = None) -> None:
        s = self.solver
        hp = s.hp
        self._halo(s.u)
        hp.stage(1, s.u, s.u, s.k1, dt)
        self._halo(s.k1)
        hp.stage(2, s.u, s.k1, s.k2, dt)
        self._halo(s.k2)
        hp.stage(3, s.u, s.k2, s.u, dt, maxabs=maxabs)
        s.launches += 3

    def solve_fixed_dt(self, dt: float | torch.Tensor, nsteps: int) -> SolveResult:
        if not isinstance(dt, torch.Tensor):
            dt = torch.full((1,), float(dt), dtype=torch.float64, device=self.solver.hp.device)
        for _ in range(nsteps):
            self.step(dt)
        return SolveResult(u=self.solver.u, steps=nsteps, t=self.solver.t)

    def solve_adaptive(self, *, theta: float, tfinal: float, cfl_scale: float, max_steps: int = 1 << 20) -> SolveResult:
        """timestepping.py:128-152 on the decomposed grid: the CFL maximum is reduced over the
        ring (one double per step), every rank then takes the identical ``dt``."""
        from . import _lib as L

        s = self.solver
        s.t.zero_()
        s.nonfinite.zero_()
        s.hp.max_abs(s.u, 1, out=s.maxabs)
        m = 0
        while m < max_steps:
            self.ring.all_max(s.maxabs)
            L.check("psk_step_control", L.lib().psk_step_control(
                1, float(theta), float(cfl_scale), float(tfinal), L.ptr(s.maxabs), L.ptr(s.t), L.ptr(s.t),
                L.ptr(s.dt), L.raw_ptr(s.active), L.raw_ptr(s.nonfinite), L.stream_ptr()))
            if int(s.active.item()) == 0:  # every rank sees the same t: uniform exit
                break
            if int(s.nonfinite.item()) != 0:
                raise ValueError("Time step is not finite.")
            s.maxabs.zero_()
            self.step(s.dt, maxabs=s.maxabs)
            m += 1
        return SolveResult(u=s.u, steps=m, t=s.t)
