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

import ctypes as ct
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
    ) -> None:
        self.hp = HotPath(
            equation=equation, flux=flux, rec=rec, bc=bc, n=n, g=g, dx=dx, eps=eps, math=math,
            nu=nu, velocity=velocity, device=device,
        )
        self.batch, self.n, self.g, self.nx = int(batch), int(n), int(g), int(n) + 2 * int(g)
        dev = self.hp.device
        # first interior cell at column 16 of a row whose stride is a multiple of 16 doubles
        self.col0 = (16 - self.g) % 16
        self.ld = ((self.col0 + self.nx + 15) // 16) * 16
        self._store = torch.zeros((3, self.batch, self.ld), dtype=torch.float64, device=dev)
        self.u, self.k1, self.k2 = (self._store[i, :, self.col0 : self.col0 + self.nx] for i in range(3))
        self.t = torch.zeros(self.batch, dtype=torch.float64, device=dev)
        self.dt = torch.zeros(self.batch, dtype=torch.float64, device=dev)
        self.maxabs = torch.zeros(self.batch, dtype=torch.float64, device=dev)
        self.active = torch.ones(self.batch, dtype=torch.uint8, device=dev)
        self.nonfinite = torch.zeros(1, dtype=torch.int32, device=dev)
        self.row_steps = torch.zeros(self.batch, dtype=torch.int32, device=dev)
        self._graph: torch.cuda.CUDAGraph | None = None
        self._graph_key: tuple | None = None
        self.launches = 0  # kernels launched by this solver (bench.py reports it)

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

    def store(self, out: torch.Tensor, *, non_blocking: bool = False) -> torch.Tensor:
        out.copy_(self.u, non_blocking=non_blocking)
        return out

    # }}}

    # {{{ one step

    def _step(self, dt: torch.Tensor, *, active: torch.Tensor | None, maxabs: torch.Tensor | None) -> None:
        hp = self.hp
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
            key = ("fixed", dt.data_ptr(), dt.numel())
            if self._graph is None or self._graph_key != key:
                # warm-up launch outside capture (lazy module loading), then capture one step
                self._store[1:].zero_()
                torch.cuda.synchronize()
                saved = self.u.clone()
                self._step(dt, active=None, maxabs=None)
                self.u.copy_(saved)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step(dt, active=None, maxabs=None)
                self.u.copy_(saved)
                self._graph, self._graph_key = g, key
            for _ in range(nsteps):
                self._graph.replay()
            self.launches += 3 * nsteps
        else:
            for _ in range(nsteps):
                self._step(dt, active=None, maxabs=None)
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
            hist = torch.zeros((min(max_steps, 1 << 16), self.batch), dtype=torch.float64, device=self.hp.device)
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
