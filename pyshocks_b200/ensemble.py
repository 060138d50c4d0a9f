"""Device-resident time loops over an ensemble of independent 1-D problems.

The reference has no batch axis; an ensemble here means "``batch`` independent runs of
the reference", one per row, each with its own time step and termination when the step is
CFL-adaptive.  State, stage buffers, time, dt and the CFL maxima all
stay on the GPU; the host only launches kernels (optionally replaying a captured CUDA
graph) and polls a termination flag every few steps.

Rows are stored with a padded stride so that the first interior cell of every row is
128-byte aligned (vectorised, fully coalesced loads/stores in the tile kernel).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from .path import HotPath


@dataclass
class SolveResult:
    u: torch.Tensor  # (batch, nx) view of the solver's state buffer
    steps: int  # largest number of steps any row took
    t: torch.Tensor  # (batch,) final times
    dt_history: np.ndarray | None = None  # (steps, batch) when recorded
    steps_per_row: np.ndarray | None = None


def row_layout(n: int, g: int) -> tuple[int, int]:
    """``(col0, ld)`` of the padded row layout: the first interior cell sits at column 16 of a row
    whose stride is a multiple of 16 doubles (128-byte aligned vector accesses)."""
    col0 = (16 - g) % 16
    return col0, ((col0 + n + 2 * g + 15) // 16) * 16


class EnsembleSolver:
    def __init__(
        self,
        *,
        equation: str,
        flux: str,
        rec: str,
        bc: str,
        n: int,
        g: int,
        dx: float,
        eps: float,
        batch: int,
        math: str = "fast",
        nu: np.ndarray | None = None,
        velocity: np.ndarray | None = None,
        device: torch.device | str | None = None,
        store: torch.Tensor | None = None,
    ) -> None:
        self.hp = HotPath(
            equation=equation, flux=flux, rec=rec, bc=bc, n=n, g=g, dx=dx, eps=eps, math=math,
            nu=nu, velocity=velocity, device=device,
        )
        self.batch, self.n, self.g, self.nx = int(batch), int(n), int(g), int(n) + 2 * int(g)
        dev = self.hp.device
        self.col0, self.ld = row_layout(self.n, self.g)
        if store is None:
            store = torch.zeros((3, self.batch, self.ld), dtype=torch.float64, device=dev)
        elif tuple(store.shape) != (3, self.batch, self.ld) or store.dtype != torch.float64:
            # caller-owned storage (peer-visible memory of the slab decomposition)
            raise ValueError(f"store must be float64 of shape {(3, self.batch, self.ld)}")
        self._store = store
        self.u, self.k1, self.k2 = self.views(self._store)
        self.t = torch.zeros(self.batch, dtype=torch.float64, device=dev)
        self.dt = torch.zeros(self.batch, dtype=torch.float64, device=dev)
        self.maxabs = torch.zeros(self.batch, dtype=torch.float64, device=dev)
        self.active = torch.ones(self.batch, dtype=torch.uint8, device=dev)
        self.nonfinite = torch.zeros(1, dtype=torch.int32, device=dev)
        self.row_steps = torch.zeros(self.batch, dtype=torch.int32, device=dev)
        self._graph: torch.cuda.CUDAGraph | None = None
        self._graph_key: tuple | None = None
        self._graph_steps = 1
        self.launches = 0  # kernels launched by this solver (bench.py reports it)
        self._fused: bool | None = None  # None: not tried yet; False: psk_ssprk33_step does not cover this scheme
        # global Lax-Friedrichs flux on periodic rows: every stage leaves max |uout| (the speed of the next stage) in
        # one of these, so no stage needs its own reduction pass (psk_ssprk33_stage_lf)
        self._lf_chain = (equation == "burgers" and flux == "lf" and bc == "periodic")
        self._lf = torch.zeros((2, 3, self.batch), dtype=torch.float64, device=dev) if self._lf_chain else None
        self._lf_parity = 0
        self._lf_ready = False  # whether _lf[1 - parity][2] holds max |u| of the current state
        self._capturing = False  # inside a CUDA-graph capture (the chain alternates buffers: not replayable as one step)

    def new_states(self, count: int) -> list[torch.Tensor]:
        """``count`` more ``(batch, nx)`` arrays with the solver's padded, aligned row layout."""
        return self.views(torch.zeros((count, self.batch, self.ld), dtype=torch.float64, device=self.hp.device))

    def views(self, store: torch.Tensor) -> list[torch.Tensor]:
        return [store[i, :, self.col0 : self.col0 + self.nx] for i in range(store.shape[0])]

    # {{{ state I/O

    def load(self, u0: torch.Tensor | np.ndarray, *, non_blocking: bool = False) -> None:
        """Copy ``(batch, nx)`` initial data (device tensor, or host array / pinned tensor)."""
        if isinstance(u0, np.ndarray):
            u0 = torch.from_numpy(np.ascontiguousarray(u0, dtype=np.float64))
        if u0.dim() == 1:
            u0 = u0[None, :]
        if tuple(u0.shape) != (self.batch, self.nx):
            raise ValueError(f"expected shape {(self.batch, self.nx)}, got {tuple(u0.shape)}")
        self.u.copy_(u0, non_blocking=non_blocking)
        self._lf_ready = False

    def store(self, out: torch.Tensor, *, non_blocking: bool = False) -> torch.Tensor:
        out.copy_(self.u, non_blocking=non_blocking)
        return out

    # }}}

    # {{{ one step

    def _step(self, dt: torch.Tensor, *, active: torch.Tensor | None, maxabs: torch.Tensor | None) -> None:
        hp = self.hp
        # hot configuration: the whole step in one launch, u -> k1, and the two arrays swap roles
        if self._fused is not False:
            self._fused = hp.step_fused(self.u, self.k1, dt, active=active, maxabs=maxabs)
            if self._fused:
                self.u, self.k1 = self.k1, self.u
                self.launches += 1
                return
        if self._lf_chain and active is None and maxabs is None and not self._capturing:
            # three stage launches + one fill, each stage taking its speed from its predecessor's maximum
            p = self._lf_parity
            if not self._lf_ready:
                hp.max_abs(self.u, 1, out=self._lf[1 - p][2])
                self.launches += 1
            cur = self._lf[p]
            cur.zero_()
            ok = hp.stage_lf(1, self.u, self.u, self.k1, dt, self._lf[1 - p][2], cur[0])
            if ok:
                hp.stage_lf(2, self.u, self.k1, self.k2, dt, cur[0], cur[1])
                hp.stage_lf(3, self.u, self.k2, self.u, dt, cur[1], cur[2])
                self._lf_parity, self._lf_ready = 1 - p, True
                self.launches += 4
                return
            self._lf_chain = False
        self._lf_ready = False
        hp.stage(1, self.u, self.u, self.k1, dt, active=active)
        hp.stage(2, self.u, self.k1, self.k2, dt, active=active)
        hp.stage(3, self.u, self.k2, self.u, dt, active=active, maxabs=maxabs)
        self.launches += 3 * (2 if (hp.equation == "burgers" and hp.flux == "lf") else 1)

    # }}}

    def solve_fixed_dt(
        self, u0: torch.Tensor | np.ndarray | None, dt: float | torch.Tensor, nsteps: int, *, graph: bool = False
    ) -> SolveResult:
        """``nsteps`` SSPRK33 steps with one shared (or per-row) fixed ``dt``
        (the fixed-step mode of the reference's convergence tests, timestepping.py:242-261)."""
        if u0 is not None:
            self.load(u0)
        if not isinstance(dt, torch.Tensor):
            dt = torch.full((1,), float(dt), dtype=torch.float64, device=self.hp.device)
        if graph and nsteps > 0:
            key = ("fixed", dt.data_ptr(), dt.numel(), self.u.data_ptr())  # (the state may sit in either array)
            if self._graph is None or self._graph_key != key:
                # warm-up launch outside capture (lazy module loading), then capture the steps.  The
                # whole-step kernel ping-pongs between two arrays, so its graph holds TWO steps
                # (u -> k1 -> u) and the captured pointers stay valid from replay to replay.
                self.k1.zero_()
                self.k2.zero_()
                torch.cuda.synchronize()
                saved = self.u.clone()
                launches = self.launches
                self._capturing = True
                self._step(dt, active=None, maxabs=None)
                per_graph = 2 if self._fused else 1
                if self._fused:
                    self.u, self.k1 = self.k1, self.u  # back to the array `saved` was taken from
                self.u.copy_(saved)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(per_graph):
                        self._step(dt, active=None, maxabs=None)
                self.u.copy_(saved)
                self._capturing = False
                self.launches = launches  # (capture and warm-up are not steps of the solve)
                self._graph, self._graph_key, self._graph_steps = g, key, per_graph
            self._lf_ready = False
            for _ in range(nsteps // self._graph_steps):
                self._graph.replay()
            self.launches += (1 if self._fused else 3) * (nsteps - nsteps % self._graph_steps)
            for _ in range(nsteps % self._graph_steps):
                self._step(dt, active=None, maxabs=None)
        else:
            for _ in range(nsteps):
                self._step(dt, active=None, maxabs=None)
        self.t += float(nsteps) * dt
        return SolveResult(u=self.u, steps=nsteps, t=self.t)

    def solve_fixed_dt_host(
        self,
        host_in: torch.Tensor,
        host_out: torch.Tensor,
        dt: float | torch.Tensor,
        nsteps: int,
        *,
        groups: int | None = None,
        streams: int = 4,
    ) -> SolveResult:
        """Host-to-host call: ``host_in`` / ``host_out`` are (pinned) host tensors of shape
        ``(batch, nx)``.  Rows are independent problems, so the batch is cut into ``groups`` row
        blocks (default: blocks of at least 1024 rows, at most 64 blocks) and each block runs
        upload -> ``nsteps`` steps -> download on its own stream: the
        PCIe copies of one block overlap the arithmetic of the others (both copy engines busy)."""
        if tuple(host_in.shape) != (self.batch, self.nx) or tuple(host_out.shape) != (self.batch, self.nx):
            raise ValueError(f"expected host tensors of shape {(self.batch, self.nx)}")
        dev = self.hp.device
        if not isinstance(dt, torch.Tensor):
            dt = torch.full((1,), float(dt), dtype=torch.float64, device=dev)
        if groups is None:
            # blocks of >= 1024 rows, at most 64 of them: the first upload and the last download are the only
            # copies not hidden behind arithmetic (measured on B200, 65536 x 4096, 20 steps: 8 blocks 7.5e10,
            # 16: 8.0e10, 32: 8.4e10, 64: 8.6e10 cell-updates/s; 4 streams beat 2 and 8)
            groups = min(64, self.batch // 1024)
        groups = max(1, min(groups, self.batch))
        if not hasattr(self, "_streams") or len(self._streams) != streams:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(streams)]
        main = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(main)
        per = (self.batch + groups - 1) // groups
        hp = self.hp
        for gi in range(groups):
            b0, b1 = gi * per, min((gi + 1) * per, self.batch)
            if b0 >= b1:
                break
            st = self._streams[gi % streams]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                u, k1, k2 = self.u[b0:b1], self.k1[b0:b1], self.k2[b0:b1]
                dtg = dt if dt.numel() == 1 else dt[b0:b1]
                u.copy_(host_in[b0:b1], non_blocking=True)
                cur, nxt = u, k1
                for _ in range(nsteps):
                    # whole step in one launch where psk_ssprk33_step covers the scheme (ping-pong)
                    if self._fused is not False:
                        self._fused = hp.step_fused(cur, nxt, dtg)
                    if self._fused:
                        cur, nxt = nxt, cur
                        self.launches += 1
                    else:
                        hp.stage(1, u, u, k1, dtg)
                        hp.stage(2, u, k1, k2, dtg)
                        hp.stage(3, u, k2, u, dtg)
                        self.launches += 3
                host_out[b0:b1].copy_(cur, non_blocking=True)
                if cur is not u:  # keep the solver's state array current for every row block alike
                    u.copy_(cur, non_blocking=True)
        for st in self._streams:
            main.wait_stream(st)
        self._lf_ready = False
        self.t += float(nsteps) * dt
        return SolveResult(u=self.u, steps=nsteps, t=self.t)

    def solve_adaptive(
        self,
        u0: torch.Tensor | np.ndarray | None,
        *,
        theta: float,
        tfinal: float,
        cfl_scale: float,
        max_steps: int = 1 << 20,
        check_every: int = 16,
        record_dt: bool = False,
    ) -> SolveResult:
        """The loop of ``timestepping.step`` (timestepping.py:128-152) per row, for Burgers-type
        schemes whose time step is ``theta * cfl_scale / max|u|`` (burgers/schemes.py:42-49,
        :121-127).  The CFL maximum is produced by stage 3 of the previous step (fused), so
        ``predict_timestep`` costs no extra pass over the state."""
        if self.hp.equation != "burgers":
            raise NotImplementedError("state-independent time steps: use solve_fixed_dt")
        if u0 is not None:
            self.load(u0)
        lib = L.lib()
        self.t.zero_()
        self.row_steps.zero_()
        self.nonfinite.zero_()
        self.hp.max_abs(self.u, 1, out=self.maxabs)
        hist = None
        if record_dt:
            max_steps = min(max_steps, 1 << 16)  # the dt history holds at most 65536 steps
            hist = torch.zeros((max_steps, self.batch), dtype=torch.float64, device=self.hp.device)
        m = 0
        while m < max_steps:
            dt_buf = self.dt if hist is None else hist[m]
            L.check(
                "psk_step_control",
                lib.psk_step_control(
                    self.batch, float(theta), float(cfl_scale), float(tfinal), L.ptr(self.maxabs), L.ptr(self.t),
                    L.ptr(self.t), L.ptr(dt_buf), L.raw_ptr(self.active), L.raw_ptr(self.nonfinite), L.stream_ptr(),
                ),
            )
            self.row_steps += self.active
            self.maxabs.zero_()
            self._step(dt_buf, active=self.active, maxabs=self.maxabs)
            self.launches += 1
            m += 1
            if m % check_every == 0 or m == max_steps:
                if int(self.nonfinite.item()) != 0:
                    raise ValueError("Time step is not finite.")  # timestepping.py:144-145
                if not bool((self.t < tfinal).any().item()):
                    break
        steps_per_row = self.row_steps.cpu().numpy()
        steps = int(steps_per_row.max())
        return SolveResult(
            u=self.u,
            steps=steps,
            t=self.t,
            dt_history=None if hist is None else hist[:steps].cpu().numpy(),
            steps_per_row=steps_per_row,
        )


class AdjointEnsemble:
    """Forward sweep with a device-resident tape and the discrete-adjoint reverse sweep for an
    ensemble advanced with a fixed ``dt`` (BASELINE.json configs[4]: B = 4096 x N = 8192, 1000
    steps, J = 1/2 sum ||u(T)||^2 over the interior).

    The reference keeps every step's state (``InMemoryCheckpoint``, timestepping.py:130-131) and
    builds a dense Jacobian per step.  Here the tape is two-level: every ``segment``-th state is
    kept (``nsteps / segment`` arrays) and the states inside a segment are recomputed into a
    ``segment``-deep scratch ring during the reverse sweep, so the memory is
    ``(nsteps / segment + segment)`` states instead of ``nsteps`` (268 MB each at config 5); each
    reverse step recomputes ``k1, k2`` (2 forward stage launches) and applies 3 fused adjoint
    stage launches.  ``p`` carries zero cotangents in the ghost cells and no boundary condition
    is imposed on it: ``backward`` returns the exact gradient of the discrete forward map."""

    def __init__(self, solver: EnsembleSolver, *, nsteps: int, dt: float | torch.Tensor,
                 segment: int | None = None, memory_fraction: float = 0.8, fused_recompute: bool = True,
                 fused_reverse: bool | None = None) -> None:
        self.s = solver
        # fused_reverse: ONE launch per reverse step (psk_ssprk33_step_adjoint: k1, k2 recomputed and the three
        # adjoint stages applied inside the kernel) wherever that kernel exists (None: automatic); the
        # stage values then never touch HBM, so the tape holds states only and the segment can be shorter
        can = solver.hp.reverse_step_supported()
        if fused_reverse and not can:
            raise ValueError("the fused reverse step needs burgers + rusanov (alpha = 1) + wenojs53, fast math, periodic or Dirichlet rows, even n")
        self.fused_reverse = can if fused_reverse is None else bool(fused_reverse)
        self.reverse_mode = ("1 launch per reverse step (psk_ssprk33_step_adjoint: recompute + 3 adjoint stages fused)"
                             if self.fused_reverse else
                             "recompute k1, k2 (1 launch) + 3 adjoint stage launches (+ 3 boundary transposes) per reverse step")
        # fused_recompute: the reverse sweep recomputes (k1, k2[, next state]) of a state with ONE launch
        # (psk_ssprk33_step_stages) instead of two or three stage launches; bit-identical stage values
        # (tests/test_gpu_adjoint_ensemble.py), 2 % off the reverse sweep of config 5.
        self.fused_recompute = bool(fused_recompute)
        self.nsteps = int(nsteps)
        dev = solver.hp.device
        if segment is None:
            segment = self._auto_segment(memory_fraction)
        self.segment = int(min(max(segment, 1), max(self.nsteps, 1)))
        self.dt = dt if isinstance(dt, torch.Tensor) else torch.full((1,), float(dt), dtype=torch.float64, device=dev)
        self.nseg = (self.nsteps + self.segment - 1) // self.segment
        self.chk = solver.new_states(self.nseg + 1)  # states at steps 0, k, 2k, ...
        # states inside the current segment and the stages k1, k2 their recomputation produces
        self.ring = solver.new_states(self.segment)
        keep_stages = 0 if self.fused_reverse else max(self.segment - 1, 0)
        self.ring_k1 = solver.new_states(keep_stages)
        self.ring_k2 = solver.new_states(keep_stages)
        if self.fused_reverse:
            self.lam1, self.p, self.pn = solver.new_states(3)
            self.lam2 = None
        else:
            self.lam2, self.lam1, self.p, self.pn = solver.new_states(4)
        self.launches = 0
        self._fused: bool | None = None  # whether psk_ssprk33_step covers the scheme (None: not tried)

    def _auto_segment(self, memory_fraction: float) -> int:
        """Smallest segment length whose tape (nsteps / k checkpoints + 3 k ring arrays) fits in
        ``memory_fraction`` of the free device memory: k = 1 keeps every state (no recompute)."""
        s = self.s
        state_bytes = s.batch * s.ld * 8
        free, _ = torch.cuda.mem_get_info(s.hp.device)
        budget = int(memory_fraction * free) // state_bytes - 8
        per_ring = 1 if self.fused_reverse else 3  # states only, or states + their stage values k1, k2
        for k in range(1, max(self.nsteps, 1) + 1):
            if (self.nsteps + k - 1) // k + 1 + per_ring * k <= budget:
                return k
        raise MemoryError("not enough device memory for the adjoint tape")

    def _advance(self, src: torch.Tensor, dst: torch.Tensor, k1: torch.Tensor | None = None,
                 k2: torch.Tensor | None = None) -> None:
        s, hp = self.s, self.s.hp
        # forward sweep (no stage values wanted): the whole step in one launch where it exists
        if k1 is None and k2 is None and self._fused is not False:
            self._fused = hp.step_fused(src, dst, self.dt)
            if self._fused:
                self.launches += 1
                return
        if k1 is not None and k2 is not None and self.fused_recompute and self._fused is not False:
            if hp.step_fused_stages(src, k1, k2, dst, self.dt):
                self.launches += 1
                return
        k1 = s.k1 if k1 is None else k1
        k2 = s.k2 if k2 is None else k2
        hp.stage(1, src, src, k1, self.dt)
        hp.stage(2, src, k1, k2, self.dt)
        hp.stage(3, src, k2, dst, self.dt)
        self.launches += 3

    def forward(self, u0: torch.Tensor | np.ndarray | None = None) -> torch.Tensor:
        """Advance ``nsteps`` steps, keeping every ``segment``-th state; returns ``u(T)`` (a view of
        the last checkpoint).  Stage 3 of a step that lands on a checkpoint writes straight into the
        checkpoint buffer, so the tape costs no extra copies."""
        s = self.s
        if u0 is not None:
            s.load(u0)
        self.chk[0].copy_(s.u)
        cur = self.chk[0]
        scratch = [s.u, self.pn]  # ping-pong for the states between two checkpoints
        for m in range(self.nsteps):
            on_chk = (m + 1) % self.segment == 0 or m + 1 == self.nsteps
            dst = self.chk[(m + self.segment) // self.segment] if on_chk else scratch[m % 2]
            self._advance(cur, dst)
            cur = dst
        return cur

    def backward(self, pT: torch.Tensor) -> torch.Tensor:
        """Reverse sweep: returns ``p(0) = (d u(T) / d u(0))^T pT`` (a view of an internal buffer)."""
        s, hp = self.s, self.s.hp
        p, pn = self.p, self.pn
        p.copy_(pT)
        if self.fused_reverse:
            for seg in range(self.nseg - 1, -1, -1):
                m0 = seg * self.segment
                last = min(m0 + self.segment, self.nsteps) - m0 - 1
                # the states u^{m0 + 1} .. u^{m0 + last} of this segment, one whole-step launch each
                for j in range(1, last + 1):
                    self._advance(self.chk[seg] if j == 1 else self.ring[j - 1], self.ring[j])
                for j in range(last, -1, -1):
                    if not hp.reverse_step_fused(self.chk[seg] if j == 0 else self.ring[j], p, self.dt, pn):
                        raise RuntimeError("psk_ssprk33_step_adjoint does not cover this configuration")
                    p, pn = pn, p
                    self.launches += 1
            self.p, self.pn = p, pn
            return p
        for seg in range(self.nseg - 1, -1, -1):
            m0 = seg * self.segment
            m1 = min(m0 + self.segment, self.nsteps)
            last = m1 - m0 - 1
            # recompute the states u^{m0} .. u^{m1 - 1} of this segment; the recomputation also
            # leaves the stages k1, k2 of every state but the last in the ring
            if self.segment == 1:
                ring0 = self.chk[seg]
            else:
                self.ring[0].copy_(self.chk[seg])
                ring0 = self.ring[0]
            for j in range(1, last + 1):
                self._advance(ring0 if j == 1 else self.ring[j - 1], self.ring[j], self.ring_k1[j - 1], self.ring_k2[j - 1])
            for j in range(last, -1, -1):
                u = ring0 if j == 0 else self.ring[j]
                if j == last:
                    k1, k2 = s.k1, s.k2
                    if self.fused_recompute and hp.step_fused_stages(u, k1, k2, None, self.dt):
                        self.launches += 1
                    else:
                        hp.stage(1, u, u, k1, self.dt)
                        hp.stage(2, u, k1, k2, self.dt)
                        self.launches += 2
                else:
                    k1, k2 = self.ring_k1[j], self.ring_k2[j]
                hp.stage_adjoint(k2, p, self.dt, 2.0 / 3.0, self.lam2)
                hp.stage_adjoint(k1, self.lam2, self.dt, 1.0 / 4.0, self.lam1)
                hp.stage_adjoint(u, self.lam1, self.dt, 1.0, pn, acc=p, c_acc=1.0 / 3.0, acc2=self.lam2, c_acc2=3.0 / 4.0)
                p, pn = pn, p
                self.launches += 3 + (3 if hp.bc in ("periodic", "neumann") else 0)
        self.p, self.pn = p, pn
        return p

    def gradient_half_l2(self, u0: torch.Tensor | np.ndarray | None = None) -> tuple[torch.Tensor, torch.Tensor]:
        """``J_b = 1/2 sum_i u_b(T)_i^2`` over the interior and ``dJ_b / du_b(0)`` for every row
        (drivers/burgers-adjoint.py:296-315: forward sweep, then the adjoint sweep started from ``p(T) = u(T)``)."""
        return self.gradient(u0)

    def gradient(self, u0: torch.Tensor | np.ndarray | None = None,
                 target: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor]:
        """``J_b = 1/2 sum_i (u_b(T)_i - target_b,i)^2`` over the interior (``target = None``: zero) and
        ``dJ_b / du_b(0)``: one forward sweep onto the tape and one reverse sweep from ``p(T) = u(T) - target``."""
        s = self.s
        uT = self.forward(u0)
        g, n = s.g, s.n
        pT = self.lam1  # scratch until the sweep starts
        pT.zero_()
        pT[:, g : g + n] = uT[:, g : g + n]
        if target is not None:
            pT[:, g : g + n] -= target[:, g : g + n]
        J = 0.5 * (pT[:, g : g + n] ** 2).sum(dim=1)
        return J, self.backward(pT)

    def optimize(self, u0: torch.Tensor | np.ndarray, *, niter: int, step: float | torch.Tensor,
                 target: torch.Tensor | None = None, callback=None) -> "OptimizeResult":
        """The optimisation loop the adjoint drivers exist for (drivers/burgers-adjoint.py:269-315 run once per
        iterate): ``niter`` steps of steepest descent on the initial condition,

            u0 <- u0 - step * dJ/du0,    J = 1/2 ||u(T; u0) - target||^2  per row,

        every iterate being one forward sweep onto the device tape and one reverse sweep (no host transfer
        inside the loop; ``J`` is read back once at the end).  ``step``: a number or one value per row.
        Returns the final initial conditions and the history of ``J`` (``niter + 1`` x batch: the last entry
        is the objective of the returned iterate)."""
        s = self.s
        s.load(u0)
        x = self.chk[0].new_empty(self.chk[0].shape)  # the iterate (forward() keeps its own copy in chk[0])
        x.copy_(s.u)
        g, n = s.g, s.n
        step_t = step if isinstance(step, torch.Tensor) else torch.full((1,), float(step), dtype=torch.float64, device=x.device)
        step_t = step_t.reshape(-1, 1)
        hist = torch.zeros((niter + 1, s.batch), dtype=torch.float64, device=x.device)
        for it in range(niter):
            J, grad = self.gradient(x, target)
            hist[it] = J
            x[:, g : g + n] -= step_t * grad[:, g : g + n]
            if callback is not None:
                callback(it, J, grad)
        uT = self.forward(x)
        r = uT[:, g : g + n] if target is None else uT[:, g : g + n] - target[:, g : g + n]
        hist[niter] = 0.5 * (r ** 2).sum(dim=1)
        return OptimizeResult(u0=x, objective=hist.cpu().numpy(), iterations=niter)


@dataclass
class OptimizeResult:
    u0: torch.Tensor  # (batch, nx) optimised initial conditions
    objective: np.ndarray  # (iterations + 1, batch) history of J
    iterations: int
