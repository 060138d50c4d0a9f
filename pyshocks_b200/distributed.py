"""Multi-GPU execution of the hot path on one 8 x B200 box: one process per GPU,
``torch.distributed`` (NCCL over NVLink / NVSwitch) for the plumbing.

Two partitionings:

* **ensemble sharding** -- rows of an ensemble are independent problems (own ``dt``, own
  termination): block-partition the batch axis, no collective in the time loop
  (:func:`shard_rows`, :class:`ShardedEnsemble`);
* **slab decomposition** of one very large periodic grid -- contiguous slabs of ``n / G`` cells
  per GPU on a ring; every RHS needs the 3 cells next to each slab edge from the neighbour
  (24 B per side per stage: latency, not bandwidth), and a CFL-adaptive ``dt`` needs one
  ``all_reduce(MAX)`` of a single double per step.  Two transports:
  :class:`PeerSlabSolver` (default for the bench) stores the edge cells straight into the
  neighbours' ghost slots through NVLink peer memory (``psk_halo_push`` / ``psk_halo_wait``,
  epoch flags, no host involvement) and overlaps the exchange with the interior cells by
  running the slab edges on a high-priority stream; :class:`SlabSolver` is the plain NCCL
  send/recv version (the baseline it is measured against, and the gloo-testable one).

The reference has no distributed path at all (platform forced to one CPU device,
``pyshocks/__init__.py:66``); these are the new workloads of BASELINE.json configs 3-5.
"""

from __future__ import annotations

from typing import Protocol, Sequence

import torch
import torch.distributed as dist

from . import _lib as L
from .ensemble import EnsembleSolver, SolveResult, row_layout
from .path import HotPath


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
        if flux == "lf":
            # the global Lax-Friedrichs speed is max|u| over the WHOLE grid (scalar.py:277): a slab only
            # sees its own cells, so every rank would use a different dissipation
            raise ValueError("slab decomposition does not support the global Lax-Friedrichs flux (use rusanov)")
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


# {{{ peer-memory transport (NVLink stores + epoch flags) with edge / interior overlap

_FLAG_WORDS = 16  # two int64 flags, padded to a 128-byte line of their own


class _RawDeviceMemory:
    """``__cuda_array_interface__`` carrier for memory torch did not allocate."""

    def __init__(self, ptr: int, words: int, typestr: str) -> None:
        self.__cuda_array_interface__ = {
            "shape": (words,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None,
        }


class PeerSlabMemory:
    """Peer-visible storage of one slab: ``u, k1, k2`` (``ld`` doubles each, the aligned row
    layout of :func:`row_layout`) followed by two epoch flags.  ``handle`` (64 bytes) lets
    another process map it (``psk_p2p_open``)."""

    def __init__(self, n_local: int, g: int, device: torch.device) -> None:
        import ctypes as ct

        self.n_local, self.g = int(n_local), int(g)
        self.col0, self.ld = row_layout(self.n_local, self.g)
        self.words = 3 * self.ld + _FLAG_WORDS
        self.device = torch.device(device)
        ptr = ct.c_void_p()
        handle = ct.create_string_buffer(64)
        with torch.cuda.device(self.device):
            L.check("psk_p2p_alloc", L.lib().psk_p2p_alloc(8 * self.words, ct.byref(ptr), handle))
        self.ptr = int(ptr.value)
        self.handle = bytes(handle.raw)
        self._raw = _RawDeviceMemory(self.ptr, self.words, "<f8")
        flat = torch.as_tensor(self._raw, device=self.device)
        self.store = flat[: 3 * self.ld].view(3, 1, self.ld)
        self._raw_flags = _RawDeviceMemory(self.ptr + 8 * 3 * self.ld, 2, "<i8")
        self.flags = torch.as_tensor(self._raw_flags, device=self.device)

    def describe(self) -> dict:
        return {"handle": self.handle, "n_local": self.n_local}

    def close(self) -> None:
        if self.ptr:
            torch.cuda.synchronize(self.device)
            self.store = self.flags = None
            L.check("psk_p2p_free", L.lib().psk_p2p_free(self.ptr))
            self.ptr = 0


class PeerRing:
    """Addresses of the two ring neighbours' ghost slots and flags, as seen from this process."""

    def __init__(self, mem: PeerSlabMemory, rank: int, world: int, left: tuple[int, int], right: tuple[int, int],
                 group: dist.ProcessGroup | None = None, opened: Sequence[int] = ()) -> None:
        self.mem, self.rank, self.world, self.group = mem, rank, world, group
        self.distributed = False  # True: the neighbours are other processes (torch.distributed is up)
        self._opened = list(opened)
        g = mem.g
        (lbase, ln), (rbase, rn) = left, right
        lcol0, lld = row_layout(ln, g)
        rcol0, rld = row_layout(rn, g)
        # array a of the LEFT neighbour: its right ghost cells; of the RIGHT neighbour: its left ghost cells
        self.dst_lo = [lbase + 8 * (a * lld + lcol0 + g + ln) for a in range(3)]
        self.dst_hi = [rbase + 8 * (a * rld + rcol0) for a in range(3)]
        # I am the left neighbour's RIGHT neighbour (its flag 1) and the right neighbour's LEFT one (flag 0)
        self.flag_lo = lbase + 8 * 3 * lld + 8
        self.flag_hi = rbase + 8 * 3 * rld
        base = mem.ptr
        self.src_lo = [base + 8 * (a * mem.ld + mem.col0 + g) for a in range(3)]
        self.src_hi = [base + 8 * (a * mem.ld + mem.col0 + mem.n_local) for a in range(3)]
        self.my_flags = (base + 8 * 3 * mem.ld, base + 8 * 3 * mem.ld + 8)

    @classmethod
    def connect(cls, mem: PeerSlabMemory, group: dist.ProcessGroup | None = None) -> "PeerRing":
        """Exchange IPC handles over ``torch.distributed`` and map the two neighbours."""
        import ctypes as ct

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        descs: list = [None] * world
        dist.all_gather_object(descs, mem.describe(), group=group)
        opened: dict[int, int] = {rank: mem.ptr}
        for r in {(rank - 1) % world, (rank + 1) % world} - {rank}:
            p = ct.c_void_p()
            with torch.cuda.device(mem.device):
                L.check("psk_p2p_open", L.lib().psk_p2p_open(descs[r]["handle"], ct.byref(p)))
            opened[r] = int(p.value)
        lr, rr = (rank - 1) % world, (rank + 1) % world
        ring = cls(mem, rank, world, (opened[lr], descs[lr]["n_local"]), (opened[rr], descs[rr]["n_local"]),
                   group=group, opened=[v for k, v in opened.items() if k != rank])
        ring.distributed = world > 1
        dist.barrier(group=group)  # every rank has mapped its neighbours before anyone pushes
        return ring

    @classmethod
    def local(cls, mems: Sequence[PeerSlabMemory], r: int) -> "PeerRing":
        """Ring between slabs held by one process (single-GPU tests of the same protocol)."""
        world = len(mems)
        lm, rm = mems[(r - 1) % world], mems[(r + 1) % world]
        return cls(mems[r], r, world, (lm.ptr, lm.n_local), (rm.ptr, rm.n_local))

    def all_max(self, x: torch.Tensor) -> None:
        if self.distributed:
            dist.all_reduce(x, op=dist.ReduceOp.MAX, group=self.group)

    def close(self) -> None:
        for p in self._opened:
            L.check("psk_p2p_close", L.lib().psk_p2p_close(p))
        self._opened = []


class PeerSlabSolver:
    """One rank's slab of a single periodic 1-D Burgers grid, ghost cells exchanged through peer
    memory and overlapped with the interior.

    Per stage ``uout = stage(u0, uin)``:

    * **edge stream** (high priority): wait until both neighbours' pushes of ``uin``'s ghost cells
      have arrived (``psk_halo_wait`` on local flags) -> stage kernel on the first and the last
      ``edge`` cells of the slab -> ``psk_halo_push`` of ``uout``'s outermost ``g`` cells into the
      neighbours' ghost slots + their flags (epoch + 1);
    * **main stream**: stage kernel on the interior ``[edge, n - edge)``, which depends only on
      local data of the previous stage, so the NVLink round trip hides behind it.

    ``fused=True`` (default whenever the configuration allows it): none of the above -- ONE launch
    per stage (``psk_ssprk33_stage_p2p``); inside the kernel only the warps that read ghost cells
    wait for the neighbours' flags, and the lanes that produce the slab's outermost cells store
    them into the neighbours' ghost slots and raise their flags.

    Events order edge(s) after interior(s - 1) and interior(s) after edge(s - 1); the time loop
    never touches the host.  A slab too short to split runs wait -> stage -> push on one stream.
    The arithmetic per cell is the same kernel on the same neighbours, so the result is
    bit-identical to the undecomposed solve."""

    def __init__(self, *, n_global: int, rank: int, world: int, dx: float, flux: str = "rusanov",
                 rec: str = "wenojs53", eps: float = 1.0e-12, math: str = "fast", edge: int = 7680,
                 overlap: bool = True, fused: bool | None = None, device: torch.device | str | None = None,
                 timeout_s: float = 20.0, whole_step: bool = False, fused_step: bool | None = None) -> None:
        if flux == "lf":
            raise ValueError("slab decomposition does not support the global Lax-Friedrichs flux (use rusanov)")
        self.rank, self.world = rank, world
        self.first, self.n_local = shard_rows(n_global, rank, world)
        self.g = g = {"constant": 1, "wenojs32": 2, "wenojs53": 3}[rec]
        # whole_step: ONE exchange of 9 cells per side and ONE launch per step (psk_ssprk33_step on a
        # slab with 9 ghost cells: three stages reach 9 cells beyond the slab) instead of three
        # exchanges of 3 cells and three launches; the state ping-pongs between arrays 0 and 1
        self.whole = bool(whole_step)
        if self.whole:
            if not (flux in ("rusanov", "godunov", "eo") and rec == "wenojs53" and math == "fast"):
                raise ValueError("whole_step needs a Burgers flux other than lf, wenojs53 and fast math")
            if fused:
                raise ValueError("whole_step and the per-stage fused exchange exclude each other")
            self.g = g = 9
            fused, overlap = False, False
        elif fused_step:
            raise ValueError("fused_step is a mode of whole_step")
        # fused_step (whole_step only; None: wherever psk_ssprk33_step_p2p covers the slab): the 9-cell exchange
        # lives INSIDE the whole-step kernel -- one launch per step instead of wait / step / push
        self.fused_step = self.whole and (fused_step is None or bool(fused_step))
        self._fused_step_required = bool(fused_step)
        self._cur = 0  # array of the store that holds the state (whole_step: 0 or 1)
        if self.n_local < 2 * g:
            raise ValueError("slabs must hold at least 2 g cells")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.mem = PeerSlabMemory(self.n_local, g, dev)
        kw = dict(equation="burgers", flux=flux, rec=rec, bc="none", g=g, dx=dx, eps=eps, math=math, device=dev)
        self.solver = EnsembleSolver(n=self.n_local, batch=1, store=self.mem.store, **kw)
        # fused: the exchange lives inside the stage kernel (psk_ssprk33_stage_p2p), one launch per
        # stage; available for the hot configuration only
        can_fuse = (flux == "rusanov" and rec == "wenojs53" and math == "fast" and self.n_local % 4 == 0
                    and self.n_local >= 8)
        if fused and not can_fuse:
            raise ValueError("the fused exchange needs rusanov + wenojs53 + fast math and n_local % 4 == 0")
        self.fused = can_fuse if fused is None else bool(fused)
        edge = (int(edge) // 16) * 16
        self.split = (not self.fused) and bool(overlap) and edge >= 16 and self.n_local >= 4 * edge
        self.edge = edge if self.split else 0
        if self.split:
            self.hp_edge = HotPath(n=self.edge, **kw)
            self.hp_mid = HotPath(n=self.n_local - 2 * self.edge, **kw)
            self.edge_stream = torch.cuda.Stream(device=dev, priority=-1)
            self.ev_main, self.ev_edge = torch.cuda.Event(), torch.cuda.Event()
        self.ring: PeerRing | None = None
        self.epoch = 0
        self.timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
        self.epoch_dev = torch.zeros(2, dtype=torch.int64, device=dev)  # device-side epoch (graph replay)
        self._graph: torch.cuda.CUDAGraph | None = None
        self._graph_key: tuple | None = None
        self.timeout_ns = int(timeout_s * 1e9)
        self.exchanges = 0
        self.launches = 0

    # {{{ set-up

    def connect(self, group: dist.ProcessGroup | None = None) -> None:
        self.ring = PeerRing.connect(self.mem, group)

    def attach(self, ring: PeerRing) -> None:
        self.ring = ring

    def close(self) -> None:
        if self.ring is not None:
            torch.cuda.synchronize(self.mem.device)
            if self.ring.distributed:
                dist.barrier(group=self.ring.group)  # nobody frees memory a neighbour still writes
            distributed, group = self.ring.distributed, self.ring.group
            self.ring.close()
            self.ring = None
            if distributed:
                dist.barrier(group=group)  # ... and nobody frees memory a neighbour still has mapped
        self.solver = None
        self.mem.close()

    @property
    def u(self) -> torch.Tensor:
        return self.solver.u

    def interior(self) -> torch.Tensor:
        return self.solver.u[0, self.g : self.g + self.n_local]

    def load_interior(self, u_local: torch.Tensor) -> None:
        """``u_local``: this rank's ``n_local`` interior cells; announces them to the neighbours."""
        self.solver.u[0, self.g : self.g + self.n_local].copy_(u_local)
        self._push(self._cur)
        if self.split:
            main = torch.cuda.current_stream()
            self.ev_main.record(main)
            self.ev_edge.record(main)

    # }}}

    # {{{ exchange primitives

    def _push(self, a: int) -> None:
        r = self.ring
        self.epoch += 1
        L.check("psk_halo_push", L.lib().psk_halo_push(
            r.src_lo[a], r.dst_lo[a], r.src_hi[a], r.dst_hi[a], self.g, r.flag_lo, r.flag_hi, self.epoch,
            L.stream_ptr()))
        self.exchanges += 1
        self.launches += 1

    def _wait(self) -> None:
        r = self.ring
        L.check("psk_halo_wait", L.lib().psk_halo_wait(
            r.my_flags[0], r.my_flags[1], self.epoch, self.timeout_ns, L.raw_ptr(self.timed_out), L.stream_ptr()))
        self.launches += 1

    def check(self) -> None:
        """Host-side check (synchronises): did a wait give up on a neighbour?"""
        if int(self.timed_out.item()) != 0:
            raise RuntimeError(f"rank {self.rank}: a neighbour's ghost cells did not arrive within {self.timeout_ns / 1e9:.0f} s")

    # }}}

    def _sub(self, a: torch.Tensor, start: int, n: int) -> torch.Tensor:
        return a[:, start : start + n + 2 * self.g]

    def _stage_fused(self, stage: int, u0: torch.Tensor, uin: torch.Tensor, uout: torch.Tensor, a_out: int,
                     dt: torch.Tensor, maxabs: torch.Tensor | None) -> None:
        import ctypes as ct

        r, hp = self.ring, self.solver.hp
        link = L.PskHaloLink()
        link.wait_lo, link.wait_hi, link.wait_epoch = r.my_flags[0], r.my_flags[1], self.epoch
        if self._dev_epoch:
            # launch arguments of a captured graph cannot change: the epoch travels in device memory,
            # stage k reads slot k & 1 and leaves epoch + 1 in slot (k + 1) & 1
            base = self.epoch_dev.data_ptr()
            link.wait_epoch = 0
            link.epoch_in = base + 8 * (self.epoch & 1)
            link.epoch_out = base + 8 * ((self.epoch + 1) & 1)
        link.peer_lo, link.peer_hi = r.dst_lo[a_out], r.dst_hi[a_out]
        link.flag_lo, link.flag_hi = r.flag_lo, r.flag_hi
        link.timeout_ns, link.timed_out = self.timeout_ns, L.raw_ptr(self.timed_out)
        batch, ld = hp._state(uin)
        d = hp.desc(batch, ld)
        L.check("psk_ssprk33_stage_p2p", L.lib().psk_ssprk33_stage_p2p(
            ct.byref(d), stage, L.ptr(u0), L.ptr(uin), L.ptr(uout), L.ptr(dt), L.ptr(maxabs), ct.byref(link),
            L.stream_ptr()))
        self.epoch += 1
        self.exchanges += 1
        self.launches += 1

    def _stage(self, stage: int, u0: torch.Tensor, uin: torch.Tensor, uout: torch.Tensor, a_out: int,
               dt: torch.Tensor, maxabs: torch.Tensor | None) -> None:
        s = self.solver
        if self.fused:
            self._stage_fused(stage, u0, uin, uout, a_out, dt, maxabs)
            return
        if not self.split:
            self._wait()
            s.hp.stage(stage, u0, uin, uout, dt, maxabs=maxabs)
            self._push(a_out)
            self.launches += 1
            return
        main = torch.cuda.current_stream()
        E, e, n = self.edge_stream, self.edge, self.n_local
        E.wait_event(self.ev_main)     # interior of the previous stage (its cells feed the edge stencils)
        main.wait_event(self.ev_edge)  # edges of the previous stage (they feed the interior stencils)
        with torch.cuda.stream(E):
            self._wait()
            for start in (0, n - e):
                self.hp_edge.stage(stage, self._sub(u0, start, e), self._sub(uin, start, e), self._sub(uout, start, e),
                                   dt, maxabs=maxabs)
            self._push(a_out)
            self.ev_edge.record(E)
        self.hp_mid.stage(stage, self._sub(u0, e, n - 2 * e), self._sub(uin, e, n - 2 * e),
                          self._sub(uout, e, n - 2 * e), dt, maxabs=maxabs)
        self.ev_main.record(main)
        self.launches += 3

    def run_stage(self, stage: int, dt: torch.Tensor, maxabs: torch.Tensor | None = None) -> None:
        """Stage 1, 2 or 3 of the current SSPRK33 step (halo wait, kernels, halo push)."""
        s = self.solver
        if stage == 1:
            if self.split:
                # dt / maxabs were produced on the main stream: the edge stream must see them
                self.ev_main.record(torch.cuda.current_stream())
            self._stage(1, s.u, s.u, s.k1, 1, dt, None)
        elif stage == 2:
            self._stage(2, s.u, s.k1, s.k2, 2, dt, None)
        else:
            self._stage(3, s.u, s.k2, s.u, 0, dt, maxabs)

    def step(self, dt: torch.Tensor, maxabs: torch.Tensor | None = None) -> None:
        if self.whole:
            # wait for the neighbours' 9 edge cells of the current state -> the whole step in one launch
            # -> push the new state's edge cells into the neighbours' ghost slots of the OTHER array
            # (they last read those slots one step ago, before the push this wait has just seen)
            s = self.solver
            if self.fused_step:
                if self._step_fused_exchange(dt, maxabs):
                    return
                if self._fused_step_required:
                    raise RuntimeError("psk_ssprk33_step_p2p does not cover this slab (shorter than 344 cells)")
                self.fused_step = False  # this slab length: wait / step / push from now on
            self._wait()
            if not s.hp.step_fused(s.u, s.k1, dt, maxabs=maxabs):
                raise RuntimeError("psk_ssprk33_step does not cover this slab configuration")
            s.u, s.k1 = s.k1, s.u
            self._cur ^= 1
            self._push(self._cur)
            self.launches += 1
            return
        for stage in (1, 2, 3):
            self.run_stage(stage, dt, maxabs)

    def _step_fused_exchange(self, dt: torch.Tensor, maxabs: torch.Tensor | None) -> bool:
        """One launch: whole step + 9-cell exchange (``psk_ssprk33_step_p2p``); ``False`` if unsupported."""
        import ctypes as ct

        s, r = self.solver, self.ring
        hp = s.hp
        nxt = self._cur ^ 1
        link = L.PskHaloLink()
        link.wait_lo, link.wait_hi, link.wait_epoch = r.my_flags[0], r.my_flags[1], self.epoch
        link.peer_lo, link.peer_hi = r.dst_lo[nxt], r.dst_hi[nxt]
        link.flag_lo, link.flag_hi = r.flag_lo, r.flag_hi
        link.timeout_ns, link.timed_out = self.timeout_ns, L.raw_ptr(self.timed_out)
        batch, ld = hp._state(s.u)
        d = hp.desc(batch, ld)
        rc = L.lib().psk_ssprk33_step_p2p(ct.byref(d), L.ptr(s.u), L.ptr(s.k1), L.ptr(dt), L.ptr(maxabs), ct.byref(link),
                                          L.stream_ptr())
        if rc == L.E_UNSUPPORTED:
            return False
        L.check("psk_ssprk33_step_p2p", rc)
        s.u, s.k1 = s.k1, s.u
        self._cur = nxt
        self.epoch += 1
        self.exchanges += 1
        self.launches += 1
        return True

    def join(self) -> None:
        """Main stream waits for the edge stream (before anything else reads the state)."""
        if self.split:
            torch.cuda.current_stream().wait_event(self.ev_edge)

    _dev_epoch = False

    def solve_fixed_dt(self, dt: float | torch.Tensor, nsteps: int, *, graph: bool = False) -> SolveResult:
        """``nsteps`` SSPRK33 steps of one shared ``dt``.  ``graph=True`` (fused exchange only) replays a
        captured CUDA graph of TWO steps (six launches, the device-side epoch alternates between two
        slots with period two stages): the host cost per step drops from three ctypes launches to half
        a graph launch, which is what small slabs are bound by."""
        if not isinstance(dt, torch.Tensor):
            dt = torch.full((1,), float(dt), dtype=torch.float64, device=self.mem.device)
        done = 0
        if graph and self.fused and nsteps >= 6:
            for _ in range(2):  # real steps first: every kernel is loaded before the capture
                self.step(dt)
            done = 2
            key = (dt.data_ptr(), self.epoch & 1)
            if self._graph is None or self._graph_key != key:
                saved = (self.epoch, self.exchanges, self.launches)
                torch.cuda.synchronize(self.mem.device)
                g = torch.cuda.CUDAGraph()
                self._dev_epoch = True
                try:
                    with torch.cuda.graph(g):
                        self.step(dt)
                        self.step(dt)
                finally:
                    self._dev_epoch = False
                self.epoch, self.exchanges, self.launches = saved  # nothing ran during the capture
                self._graph, self._graph_key = g, key
            replays = (nsteps - done) // 2
            self.epoch_dev[self.epoch & 1].fill_(self.epoch)
            for _ in range(replays):
                self._graph.replay()
            self.epoch += 6 * replays
            self.exchanges += 6 * replays
            self.launches += 6 * replays
            done += 2 * replays
        for _ in range(nsteps - done):
            self.step(dt)
        self.join()
        self.check()  # a ghost-cell wait that gave up would otherwise yield a silently wrong state
        return SolveResult(u=self.solver.u, steps=nsteps, t=self.solver.t)

    def solve_adaptive(self, *, theta: float, tfinal: float, cfl_scale: float, max_steps: int = 1 << 20) -> SolveResult:
        """timestepping.py:128-152 on the decomposed grid (see :meth:`SlabSolver.solve_adaptive`)."""
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
            if int(s.active.item()) == 0:
                break
            if int(s.nonfinite.item()) != 0:
                raise ValueError("Time step is not finite.")
            s.maxabs.zero_()
            self.step(s.dt, maxabs=s.maxabs)
            self.join()
            m += 1
        self.check()
        return SolveResult(u=s.u, steps=m, t=s.t)


# }}}


# {{{ discrete adjoint on slabs


class PeerSlabAdjoint:
    """Forward sweep with a device tape and the discrete-adjoint reverse sweep of ONE periodic Burgers grid
    (WENO-JS5 + Rusanov + SSPRK33, fixed ``dt``) that is slab-decomposed over the ranks -- the adjoint of
    BASELINE config 4 ("adjoint on slabs").

    Every array of a rank carries 16 ghost cells per side in peer-visible memory:

    * forward: one whole-step launch per step (``psk_ssprk33_step``, boundary kind NONE) writes state ``m + 1``
      straight onto the tape, and ``psk_halo_push`` stores its 16 edge cells into the neighbours' ghost slots
      of THEIR tape entry ``m + 1`` -- so every checkpoint already holds the neighbours' cells the reverse
      sweep will need;
    * reverse: one launch per step (``psk_ssprk33_step_adjoint``, boundary kind NONE) GATHERS the slab's part of
      ``p^m = (d u^{m+1} / d u^m)^T p^{m+1}`` from ``u^m`` and ``p^{m+1}`` of the slab plus 16 cells of either
      neighbour; then the 16 edge cells of ``p^m`` go to the neighbours.  SURVEY.md 8(e) sketched the scatter
      form (ghost contributions "sent back and added" to the owner); recomputing the 16-cell overlap instead
      needs one exchange per step in the same direction as the forward one and no atomics or adds.

    Epoch flags order the exchange exactly as in :class:`PeerSlabSolver` (``psk_halo_wait`` before a step,
    ``psk_halo_push`` after it); a neighbour's ghost slots are only overwritten after its push of the step that
    read them has been observed."""

    G = 16

    def __init__(self, *, n_global: int, rank: int, world: int, dx: float, nsteps: int, eps: float = 1.0e-12,
                 device: torch.device | str | None = None, timeout_s: float = 20.0) -> None:
        import ctypes as ct

        if n_global % (2 * world) != 0:
            raise ValueError("the slab adjoint needs equal slabs of even length (n_global % (2 world) == 0)")
        self.rank, self.world, self.nsteps = rank, world, int(nsteps)
        self.first, self.n_local = shard_rows(n_global, rank, world)
        g = self.G
        if self.n_local < 2 * g:
            raise ValueError("slabs must hold at least 32 cells")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        self.col0, self.ld = row_layout(self.n_local, g)
        self.narrays = self.nsteps + 3  # tape 0 .. nsteps, p ping, p pong
        self.words = self.narrays * self.ld + _FLAG_WORDS
        ptr = ct.c_void_p()
        handle = ct.create_string_buffer(64)
        with torch.cuda.device(dev):
            L.check("psk_p2p_alloc", L.lib().psk_p2p_alloc(8 * self.words, ct.byref(ptr), handle))
        self.ptr, self.handle = int(ptr.value), bytes(handle.raw)
        self._raw = _RawDeviceMemory(self.ptr, self.words, "<f8")
        flat = torch.as_tensor(self._raw, device=dev)
        flat.zero_()
        self.nx = self.n_local + 2 * g
        self.arrays = flat[: self.narrays * self.ld].view(self.narrays, 1, self.ld)[:, :, self.col0 : self.col0 + self.nx]
        self.hp = HotPath(equation="burgers", flux="rusanov", rec="wenojs53", bc="none", n=self.n_local, g=g, dx=dx,
                          eps=eps, math="fast", device=dev)
        self.timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
        self.timeout_ns = int(timeout_s * 1e9)
        self.epoch = 0
        self.launches = 0
        self._peers: tuple[int, int] | None = None  # base addresses of the left / right neighbour
        self._opened: list[int] = []
        self._group = None
        self._distributed = False

    # {{{ set-up

    def connect(self, group: dist.ProcessGroup | None = None) -> None:
        import ctypes as ct

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        descs: list = [None] * world
        dist.all_gather_object(descs, {"handle": self.handle, "n_local": self.n_local}, group=group)
        opened: dict[int, int] = {rank: self.ptr}
        for r in {(rank - 1) % world, (rank + 1) % world} - {rank}:
            p = ct.c_void_p()
            with torch.cuda.device(self.device):
                L.check("psk_p2p_open", L.lib().psk_p2p_open(descs[r]["handle"], ct.byref(p)))
            opened[r] = int(p.value)
        self._peers = (opened[(rank - 1) % world], opened[(rank + 1) % world])
        self._opened = [v for k, v in opened.items() if k != rank]
        self._group, self._distributed = group, world > 1
        dist.barrier(group=group)

    def attach_local(self, slabs: Sequence["PeerSlabAdjoint"]) -> None:
        """Ring between slabs held by one process (single-GPU tests of the same protocol)."""
        w = len(slabs)
        self._peers = (slabs[(self.rank - 1) % w].ptr, slabs[(self.rank + 1) % w].ptr)

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        if self._distributed:
            dist.barrier(group=self._group)
        for p in self._opened:
            L.check("psk_p2p_close", L.lib().psk_p2p_close(p))
        self._opened = []
        if self._distributed:
            dist.barrier(group=self._group)
        self.arrays = None
        if self.ptr:
            L.check("psk_p2p_free", L.lib().psk_p2p_free(self.ptr))
            self.ptr = 0

    # }}}

    def _addr(self, base: int, k: int, cell: int) -> int:
        """address of interior cell ``cell`` (may be negative / beyond: ghost cells) of array ``k``"""
        return base + 8 * (k * self.ld + self.col0 + self.G + cell)

    def _push(self, k: int) -> None:
        """my 16 edge cells of array k -> the neighbours' ghost slots of THEIR array k, then their flags"""
        lbase, rbase = self._peers
        g, n = self.G, self.n_local
        flags = 8 * self.narrays * self.ld
        self.epoch += 1
        L.check("psk_halo_push", L.lib().psk_halo_push(
            self._addr(self.ptr, k, 0), self._addr(lbase, k, n),       # my first cells -> left neighbour's right ghosts
            self._addr(self.ptr, k, n - g), self._addr(rbase, k, -g),  # my last cells -> right neighbour's left ghosts
            g, lbase + flags + 8, rbase + flags, self.epoch, L.stream_ptr()))
        self.launches += 1

    def _wait(self) -> None:
        flags = self.ptr + 8 * self.narrays * self.ld
        L.check("psk_halo_wait", L.lib().psk_halo_wait(flags, flags + 8, self.epoch, self.timeout_ns,
                                                      L.raw_ptr(self.timed_out), L.stream_ptr()))
        self.launches += 1

    def check(self) -> None:
        if int(self.timed_out.item()) != 0:
            raise RuntimeError(f"rank {self.rank}: a neighbour's ghost cells did not arrive within {self.timeout_ns / 1e9:.0f} s")

    def interior(self, k: int) -> torch.Tensor:
        return self.arrays[k, 0, self.G : self.G + self.n_local]

    # {{{ sweeps (split into phases so that in-process rings can interleave the slabs)

    def forward_begin(self, u0_local: torch.Tensor) -> None:
        self.interior(0).copy_(u0_local)
        self._push(0)

    def forward_step(self, m: int, dt: torch.Tensor) -> None:
        self._wait()
        if not self.hp.step_fused(self.arrays[m], self.arrays[m + 1], dt):
            raise RuntimeError("psk_ssprk33_step does not cover this slab")
        self._push(m + 1)
        self.launches += 1

    def backward_begin(self, pT_local: torch.Tensor) -> None:
        self._cur = self.nsteps + 1
        self.interior(self._cur).copy_(pT_local)
        self._push(self._cur)

    def backward_step(self, m: int, dt: torch.Tensor) -> None:
        cur = self._cur
        nxt = 2 * self.nsteps + 3 - cur  # the other of nsteps + 1, nsteps + 2
        self._wait()
        if not self.hp.reverse_step_fused(self.arrays[m], self.arrays[cur], dt, self.arrays[nxt]):
            raise RuntimeError("psk_ssprk33_step_adjoint does not cover this slab")
        self._push(nxt)
        self._cur = nxt
        self.launches += 1

    def gradient_half_l2(self, u0_local: torch.Tensor, dt: float | torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """``J = 1/2 sum u(T)^2`` over this slab and this slab's part of ``dJ / du(0)`` (forward sweep onto the tape,
        reverse sweep from ``p(T) = u(T)``); every rank calls it at the same time."""
        if not isinstance(dt, torch.Tensor):
            dt = torch.full((1,), float(dt), dtype=torch.float64, device=self.device)
        self.forward_begin(u0_local)
        for m in range(self.nsteps):
            self.forward_step(m, dt)
        uT = self.interior(self.nsteps)
        J = 0.5 * (uT ** 2).sum()
        self.backward_begin(uT)
        for m in range(self.nsteps - 1, -1, -1):
            self.backward_step(m, dt)
        self.check()
        return J, self.interior(self._cur)

    # }}}


# }}}

