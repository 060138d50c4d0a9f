"""`HotPath`: one bound (equation, flux, reconstruction, boundary kind, grid) on the GPU.

This is the layer right above the C ABI: it owns the small device-side tables a
scheme needs (per-face dissipation ``nu``, velocity and its reconstruction, ghost
data), builds the ``psk_desc`` for each call and launches the kernels of
``libpsk.so`` on the current torch CUDA stream.  The pyshocks-shaped API
(``apply_operator``, ``advance``, ...) in the sibling modules dispatches here.

All state arrays are ``torch.float64`` CUDA tensors of shape ``(nx,)`` or
``(batch, nx)`` with ``nx = n + 2 g`` (ghost cells included, as in the reference).
"""

from __future__ import annotations

import ctypes as ct
from typing import Sequence

import numpy as np
import torch

from . import _lib as L

_EQ = {"burgers": L.EQ_BURGERS, "advection": L.EQ_ADVECTION, "continuity": L.EQ_CONTINUITY}
_FLUX = {
    "rusanov": L.FLUX_RUSANOV,
    "lf": L.FLUX_LAX_FRIEDRICHS,
    "godunov": L.FLUX_UPWIND,
    "upwind": L.FLUX_UPWIND,
    "eo": L.FLUX_ENGQUIST_OSHER,
    "esweno32": L.FLUX_ESWENO,
}
_REC = {"constant": L.REC_CONSTANT, "wenojs32": L.REC_WENOJS32, "wenojs53": L.REC_WENOJS53,
        "esweno32": L.REC_ESWENO32}
_BC = {"periodic": L.BC_PERIODIC, "dirichlet": L.BC_DIRICHLET, "neumann": L.BC_NEUMANN, "none": L.BC_NONE}
_MATH = {"fast": L.MATH_FAST, "strict": L.MATH_STRICT}


def _dev(device: torch.device | str | None) -> torch.device:
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("pyshocks_b200 needs a CUDA device (B200); there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _like(u: torch.Tensor) -> torch.Tensor:
    """uninitialised array with the shape AND row stride of ``u`` (views of padded storage
    keep their stride, so all arrays of one launch share ``ld``)"""
    return torch.empty_strided(u.shape, u.stride(), dtype=u.dtype, device=u.device)


class HotPath:
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
        math: str = "fast",
        nu: np.ndarray | torch.Tensor | None = None,
        velocity: np.ndarray | torch.Tensor | None = None,
        device: torch.device | str | None = None,
        delta: float = 0.0,
    ) -> None:
        self.device = _dev(device)
        self.delta = float(delta)  # ESWENO32 only
        self.equation, self.flux, self.rec, self.bc, self.math = equation, flux, rec, bc, math
        self.n, self.g, self.nx = int(n), int(g), int(n) + 2 * int(g)
        self.dx, self.eps = float(dx), float(eps)
        self._nu = None if nu is None else self._table(nu)
        self._vel = self._vel_l = self._vel_r = None
        self._vel_periodic = False
        self._ghost: torch.Tensor | None = None
        self._ghost3: torch.Tensor | None = None
        self._ghost_ld = 0
        self._work: dict[tuple, torch.Tensor] = {}
        self._ends: HotPath | None = None
        if velocity is not None:
            self._vel = self._table(velocity)
            if self._vel.numel() != self.nx:
                raise ValueError("velocity must have nx = n + 2 g entries")
            # reconstruct(rec, grid, bc, a, a, a) once: velocity does not depend on time
            # (advection/schemes.py:104-105, continuity/schemes.py:100-101)
            self._vel_l, self._vel_r = self.reconstruct(self._vel)
            # whether the velocity's reconstruction is periodic like the state (the condition of psk_ssprk33_step
            # for advection / continuity on periodic rows); false for a velocity array whose ghost cells are not the
            # images of its interior: three stage launches then
            g, n = self.g, self.n
            self._vel_periodic = bool(self._vel_r[g - 1] == self._vel_r[g + n - 1]) and bool(self._vel_l[g + n] == self._vel_l[g])
        elif equation != "burgers":
            raise ValueError(f"{equation} schemes need a velocity array")

    # {{{ helpers

    def _table(self, a: np.ndarray | torch.Tensor) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            return a.detach().to(device=self.device, dtype=torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)

    def workspace(self, key: str, shape: Sequence[int], dtype: torch.dtype = torch.float64) -> torch.Tensor:
        k = (key, tuple(shape), dtype)
        buf = self._work.get(k)
        if buf is None:
            buf = torch.zeros(tuple(shape), dtype=dtype, device=self.device)
            self._work[k] = buf
        return buf

    def set_ghost(self, values: np.ndarray | torch.Tensor | None) -> None:
        """Dirichlet values / Neumann offsets for the next calls: ``(2g,)`` shared by all
        rows or ``(batch, 2g)``; left ghosts first (include/psk.h, enum psk_bc)."""
        self._ghost3 = None
        if values is None:
            self._ghost, self._ghost_ld = None, 0
            return
        gh = self._table(values)
        if gh.shape[-1] != 2 * self.g:
            raise ValueError(f"ghost data needs 2 g = {2 * self.g} values per row")
        self._ghost = gh
        self._ghost_ld = 0 if gh.dim() == 1 else gh.stride(0)

    def ghost3(self, ghosts: Sequence[np.ndarray | torch.Tensor] | torch.Tensor | None = None) -> torch.Tensor | None:
        """Boundary data of the three stage times as ONE device array ``(3, 2 g)`` or ``(3, batch, 2 g)`` (what
        ``psk_ssprk33_step_bc`` takes); ``None``: the data of :meth:`set_ghost`, the same at all three times."""
        if ghosts is None:
            if self._ghost is None:
                return None
            if self._ghost3 is None:
                self._ghost3 = torch.stack([self._ghost.contiguous()] * 3).contiguous()
            return self._ghost3
        if isinstance(ghosts, torch.Tensor):
            g3 = ghosts.to(device=self.device, dtype=torch.float64).contiguous()
        else:
            g3 = torch.stack([self._table(x) for x in ghosts]).contiguous()
        if g3.shape[0] != 3 or g3.shape[-1] != 2 * self.g or g3.dim() not in (2, 3):
            raise ValueError(f"ghost data of a step: three sets of 2 g = {2 * self.g} values (per row)")
        return g3

    def desc(self, batch: int, ld: int, *, bc: int | None = None) -> L.PskDesc:
        d = L.PskDesc()
        d.equation, d.flux, d.rec = _EQ[self.equation], _FLUX[self.flux], _REC[self.rec]
        d.bc = _BC[self.bc] if bc is None else bc
        d.math = _MATH[self.math]
        d.n, d.g, d.batch, d.ld = self.n, self.g, batch, ld
        d.dx, d.eps, d.delta = self.dx, self.eps, self.delta
        d.nu = L.ptr(self._nu)
        d.velocity, d.vel_l, d.vel_r = L.ptr(self._vel), L.ptr(self._vel_l), L.ptr(self._vel_r)
        d.ghost, d.ghost_ld = L.ptr(self._ghost), self._ghost_ld
        return d

    def _state(self, u: torch.Tensor) -> tuple[int, int]:
        batch, nx, ld = L.rows_of(u)
        if nx != self.nx:
            raise ValueError(f"array has {nx} cells per row, grid has nx = {self.nx}")
        L.ptr(u)
        if self._ghost is not None and self._ghost.dim() == 2 and self._ghost.shape[0] != batch:
            raise ValueError("per-row ghost data does not match the batch size")
        return batch, ld

    def _lf_work(self, batch: int) -> torch.Tensor | None:
        # scratch is per launching stream: row blocks of one solver run concurrently on several streams
        # (EnsembleSolver.solve_fixed_dt_host) and must not share the speed buffer
        if self.equation == "burgers" and self.flux == "lf":
            return self.workspace(f"lf@{torch.cuda.current_stream(self.device).cuda_stream}", (batch,))
        return None

    # }}}

    # {{{ parity entry points

    def apply_boundary(self, u: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        batch, ld = self._state(u)
        w = _like(u) if out is None else out
        if L.rows_of(w)[2] != ld:
            raise ValueError("output must share the row stride of the input")
        d = self.desc(batch, ld)
        L.check("psk_apply_boundary", L.lib().psk_apply_boundary(ct.byref(d), L.ptr(u), L.ptr(w), L.stream_ptr()))
        return w

    def reconstruct(self, f: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        batch, nx, ld = L.rows_of(f)
        if nx != self.nx:
            raise ValueError(f"array has {nx} cells per row, grid has nx = {self.nx}")
        fl, fr = torch.empty_like(f), torch.empty_like(f)
        if L.rows_of(fl)[2] != ld:
            f = f.contiguous()
            fl, fr = torch.empty_like(f), torch.empty_like(f)
            ld = L.rows_of(f)[2]
        d = self.desc(batch, ld)
        L.check("psk_reconstruct", L.lib().psk_reconstruct(ct.byref(d), L.ptr(f), L.ptr(fl), L.ptr(fr), L.stream_ptr()))
        return fl, fr

    def numerical_flux(self, w: torch.Tensor) -> torch.Tensor:
        batch, ld = self._state(w)
        shape = (self.nx + 1,) if w.dim() == 1 else (batch, self.nx + 1)
        F = torch.empty(shape, dtype=torch.float64, device=w.device)
        d = self.desc(batch, ld, bc=L.BC_NONE)  # w already carries its boundary data
        L.check(
            "psk_numerical_flux",
            L.lib().psk_numerical_flux(
                ct.byref(d), L.ptr(w), L.ptr(F), self.nx + 1, L.ptr(self._lf_work(batch)), L.stream_ptr()
            ),
        )
        return F

    def apply_operator(self, u: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        batch, ld = self._state(u)
        rhs = _like(u) if out is None else out
        if L.rows_of(rhs)[2] != ld:
            raise ValueError("output must share the row stride of the input")
        d = self.desc(batch, ld)
        L.check(
            "psk_apply_operator",
            L.lib().psk_apply_operator(ct.byref(d), L.ptr(u), L.ptr(rhs), L.ptr(self._lf_work(batch)), L.stream_ptr()),
        )
        return rhs

    def max_abs(self, u: torch.Tensor, mode: int = 1, out: torch.Tensor | None = None) -> torch.Tensor:
        batch, ld = self._state(u)
        res = torch.empty((batch,), dtype=torch.float64, device=u.device) if out is None else out
        d = self.desc(batch, ld)
        L.check("psk_max_abs", L.lib().psk_max_abs(ct.byref(d), L.ptr(u), mode, L.ptr(res), L.stream_ptr()))
        return res

    # }}}

    # {{{ discrete adjoint

    def _adj_work(self, batch: int) -> torch.Tensor:
        return self.workspace(f"adj@{torch.cuda.current_stream(self.device).cuda_stream}", (batch * (2 * self.g + 3),))

    def apply_operator_vjp(self, u: torch.Tensor, v: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """``J_L(u)^T v`` for ``L = apply_operator`` (what ``jax.vjp(apply_operator)`` returns)."""
        batch, ld = self._state(u)
        res = _like(u) if out is None else out
        if L.rows_of(v)[2] != ld or L.rows_of(res)[2] != ld:
            raise ValueError("all arrays must share one row stride")
        d = self.desc(batch, ld)
        L.check(
            "psk_apply_operator_vjp",
            L.lib().psk_apply_operator_vjp(
                ct.byref(d), L.ptr(u), L.ptr(v), L.ptr(res), L.ptr(self._adj_work(batch)), L.stream_ptr()
            ),
        )
        return res

    def stage_adjoint(
        self,
        x: torch.Tensor,
        v: torch.Tensor,
        dt: torch.Tensor,
        c_v: float,
        out: torch.Tensor,
        *,
        acc: torch.Tensor | None = None,
        c_acc: float = 0.0,
        acc2: torch.Tensor | None = None,
        c_acc2: float = 0.0,
    ) -> torch.Tensor:
        """``out = c_acc acc + c_acc2 acc2 + c_v (v + dt J_L(x)^T v)`` in one launch."""
        batch, ld = self._state(x)
        for a in (v, out, acc, acc2):
            if a is not None and L.rows_of(a)[2] != ld:
                raise ValueError("all arrays must share one row stride")
        d = self.desc(batch, ld)
        L.check(
            "psk_ssprk33_stage_adjoint",
            L.lib().psk_ssprk33_stage_adjoint(
                ct.byref(d), L.ptr(x), L.ptr(v), L.ptr(dt), 0 if dt.numel() == 1 else 1, float(c_v),
                L.ptr(acc), float(c_acc), L.ptr(acc2), float(c_acc2), L.ptr(self._adj_work(batch)),
                L.ptr(out), L.stream_ptr(),
            ),
        )
        return out

    def ssprk33_step_adjoint(
        self,
        u: torch.Tensor,
        dt: torch.Tensor,
        p: torch.Tensor,
        *,
        out: torch.Tensor | None = None,
        ghosts: Sequence[np.ndarray | torch.Tensor] | None = None,
        stages: tuple[torch.Tensor, torch.Tensor] | None = None,
    ) -> torch.Tensor:
        """``(d advance / d u)^T p`` for one SSPRK33 step from the checkpointed state ``u``
        (timestepping.py:205-206 without the dense Jacobian): recompute ``k1, k2``, then three
        fused adjoint stages."""
        if stages is None:
            k1, k2 = _like(u), _like(u)
            if ghosts is not None:
                self.set_ghost(ghosts[0])
            self.stage(1, u, u, k1, dt, ghost_rows=True)
            if ghosts is not None:
                self.set_ghost(ghosts[1])
            self.stage(2, u, k1, k2, dt, ghost_rows=True)
        else:
            k1, k2 = stages
        lam2, lam1 = _like(u), _like(u)
        res = _like(u) if out is None else out
        if ghosts is not None:
            self.set_ghost(ghosts[2])
        self.stage_adjoint(k2, p, dt, 2.0 / 3.0, lam2)
        if ghosts is not None:
            self.set_ghost(ghosts[1])
        self.stage_adjoint(k1, lam2, dt, 1.0 / 4.0, lam1)
        if ghosts is not None:
            self.set_ghost(ghosts[0])
        self.stage_adjoint(u, lam1, dt, 1.0, res, acc=p, c_acc=1.0 / 3.0, acc2=lam2, c_acc2=3.0 / 4.0)
        return res

    # }}}

    # {{{ whole solve in one launch (small rows)

    def solve_rows(
        self,
        u: torch.Tensor,
        *,
        tfinal: float | None = None,
        theta: float = 1.0,
        cfl_scale: float | None = None,
        fixed_dt: float | None = None,
        max_steps: int,
        record_dt: bool = False,
        tape: bool = False,
    ) -> dict:
        """Advance ``u`` IN PLACE with the whole ``timestepping.step`` loop in one kernel launch
        (one CTA per row, state in shared memory).  Adaptive (Burgers, ``cfl_scale`` given) or
        ``max_steps`` steps of ``fixed_dt``.  Returns ``{"t", "steps", "dt", "tape"}`` (device tensors)."""
        batch, ld = self._state(u)
        adaptive = fixed_dt is None
        if adaptive and (tfinal is None or cfl_scale is None):
            raise ValueError("adaptive solves need tfinal and cfl_scale")
        dev = u.device
        t_out = torch.zeros(batch, dtype=torch.float64, device=dev)
        steps = torch.zeros(batch, dtype=torch.int32, device=dev)
        hist = torch.zeros((batch, max_steps), dtype=torch.float64, device=dev) if record_dt else None
        tp = None
        if tape:
            tp = torch.zeros((max_steps + 1, batch, ld), dtype=torch.float64, device=dev)
        d = self.desc(batch, ld)
        L.check(
            "psk_solve_rows",
            L.lib().psk_solve_rows(
                ct.byref(d), L.ptr(u), int(adaptive), float(theta), float(cfl_scale or 0.0),
                float(tfinal if tfinal is not None else 0.0), float(fixed_dt or 0.0), int(max_steps),
                L.ptr(t_out), L.raw_ptr(steps), L.ptr(hist), L.ptr(tp), L.stream_ptr(),
            ),
        )
        return {"t": t_out, "steps": steps, "dt": hist, "tape": None if tp is None else tp[:, :, : self.nx]}

    def solve_rows_tables(self, u: torch.Tensor, dts: torch.Tensor, ghost_table: torch.Tensor | None = None, *,
                          tape: bool = False) -> dict:
        """:meth:`solve_rows` with one step size per step (``dts``: device ``(nsteps,)``) and, for boundary data
        that depend on time, ``ghost_table``: device ``(nsteps, 3, 2 g)`` -- the Dirichlet values / Neumann
        offsets at the stage times ``t, t + dt, t + dt / 2`` of every step (``psk_solve_rows_tables``)."""
        batch, ld = self._state(u)
        nsteps = int(dts.numel())
        dev = u.device
        t_out = torch.zeros(batch, dtype=torch.float64, device=dev)
        steps = torch.zeros(batch, dtype=torch.int32, device=dev)
        tp = torch.zeros((nsteps + 1, batch, ld), dtype=torch.float64, device=dev) if tape else None
        if ghost_table is not None:
            if tuple(ghost_table.shape) != (nsteps, 3, 2 * self.g) or not ghost_table.is_contiguous():
                raise ValueError(f"ghost_table must be contiguous of shape {(nsteps, 3, 2 * self.g)}")
        d = self.desc(batch, ld)
        if ghost_table is not None:
            d.ghost, d.ghost_ld = L.ptr(ghost_table), 0  # (validated as "present"; the kernel takes the table)
        L.check(
            "psk_solve_rows_tables",
            L.lib().psk_solve_rows_tables(ct.byref(d), L.ptr(u), nsteps, L.ptr(dts), L.ptr(ghost_table), L.ptr(t_out),
                                          L.raw_ptr(steps), L.ptr(tp), L.stream_ptr()),
        )
        return {"t": t_out, "steps": steps, "tape": tp}

    def steps_tape(self, u0: torch.Tensor, dts: torch.Tensor, ghost_table: torch.Tensor | None = None) -> torch.Tensor | None:
        """``nsteps`` whole-step launches from one call (``psk_ssprk33_steps_tape``), every state written straight
        onto a freshly allocated tape with the aligned row layout: returns the ``(nsteps + 1, batch, nx)`` view of
        the states (``[0]`` = ``u0``, ``[-1]`` = the final state), or ``None`` -- nothing launched -- where the
        whole-step kernel does not exist."""
        from .ensemble import row_layout

        if u0.dim() == 1:
            u0 = u0[None, :]
        batch, nx = u0.shape
        if nx != self.nx:
            raise ValueError(f"array has {nx} cells per row, grid has nx = {self.nx}")
        nsteps = int(dts.numel())
        col0, ld = row_layout(self.n, self.g)
        # (zeros: the whole-step kernel writes interior cells only; stored ghost cells are never read)
        store = torch.zeros((nsteps + 1, batch, ld), dtype=torch.float64, device=u0.device)
        tape = store[:, :, col0 : col0 + nx]
        tape[0].copy_(u0)
        d = self.desc(batch, ld)
        if ghost_table is not None:
            if tuple(ghost_table.shape) != (nsteps, 3, 2 * self.g) or not ghost_table.is_contiguous():
                raise ValueError(f"ghost_table must be contiguous of shape {(nsteps, 3, 2 * self.g)}")
            d.ghost, d.ghost_ld = L.ptr(ghost_table), 0
        rc = L.lib().psk_ssprk33_steps_tape(ct.byref(d), L.ptr(tape), batch * ld, nsteps, L.ptr(dts), L.ptr(ghost_table),
                                            L.stream_ptr())
        if rc == L.E_UNSUPPORTED:
            return None
        L.check("psk_ssprk33_steps_tape", rc)
        return tape

    def adjoint_sweep(self, tape: torch.Tensor, dts: torch.Tensor, p: torch.Tensor, *,
                      ghost_table: torch.Tensor | None = None, p_boundary: "HotPath | None" = None,
                      history: bool = False) -> torch.Tensor | None:
        """The whole loop of ``adjoint_step`` (timestepping.py:198-209) in one call (``psk_ssprk33_adjoint_sweep``):
        ``p`` (``(batch, nx)`` with the tape's row stride, already passed through the adjoint boundary condition)
        is replaced by ``p(0)``.  ``tape``: ``(nsteps + 1, batch, ld)`` from :meth:`solve_rows` /
        :meth:`solve_rows_tables`; ``p_boundary``: the binding whose boundary kind (and ghost data) is imposed on
        ``p`` after every step.  Returns every intermediate ``p`` (``(nsteps, batch, ld)``) if ``history``."""
        nsteps = int(dts.numel())
        if tape.dim() != 3 or tape.shape[0] < nsteps + 1 or tape.shape[2] != self.nx or tape.stride(2) != 1:
            raise ValueError("tape must be (nsteps + 1, batch, nx) with unit stride along x")
        batch, ld = tape.shape[1], (tape.stride(1) if tape.shape[1] > 1 else max(tape.stride(1), self.nx))
        pb, pnx, pld = L.rows_of(p)
        if (pb, pnx) != (batch, self.nx) or (pb > 1 and pld != ld):
            raise ValueError("p must be a (batch, nx) view with the row stride of the tape")
        dev = tape.device
        # scratch with the row layout of the tape (same stride, same offset of the first cell modulo 16 bytes)
        lead = (tape.data_ptr() // 8) % 2
        flat = torch.empty((5 * batch * ld + lead + 2,), dtype=torch.float64, device=dev)
        off = (lead - (flat.data_ptr() // 8) % 2) % 2
        states = flat[off : off + 5 * batch * ld]
        hist = torch.empty((nsteps, batch, ld), dtype=torch.float64, device=dev) if history else None
        d = self.desc(batch, ld)
        if ghost_table is not None:
            d.ghost, d.ghost_ld = L.ptr(ghost_table), 0
        pd = None
        if p_boundary is not None:
            pd = p_boundary.desc(batch, ld)
            # Homogeneous Dirichlet data on the adjoint variable (drivers/advection-adjoint.py): once the ghost cells of
            # p are zero they stay zero -- every transposed stage writes c_v v + c_acc acc + c_acc2 acc2 there, the
            # cotangents that land on ghost cells go to the cells they were copied from or nowhere -- so
            # apply_boundary(p) after a step changes nothing and its launch (one in five of a reverse step) is dropped
            g = self.g
            if (p_boundary.bc == "dirichlet" and self.bc != "none" and p_boundary._ghost is not None
                    and not bool(p_boundary._ghost.any())
                    and not bool(p[:, :g].any()) and not bool(p[:, self.nx - g : self.nx].any())):
                pd = None
        L.check(
            "psk_ssprk33_adjoint_sweep",
            L.lib().psk_ssprk33_adjoint_sweep(
                ct.byref(d), L.ptr(tape), tape.stride(0), nsteps, L.ptr(dts), L.ptr(ghost_table),
                ct.byref(pd) if pd is not None else None, L.ptr(p), L.ptr(states), L.ptr(self._adj_work(batch)),
                L.ptr(self._lf_work(batch)), L.ptr(hist), L.stream_ptr()),
        )
        return None if hist is None else hist[:, :, : self.nx]

    # }}}

    # {{{ fused SSPRK33

    def stage(
        self,
        stage: int,
        u0: torch.Tensor | None,
        uin: torch.Tensor,
        uout: torch.Tensor,
        dt: torch.Tensor | None,
        *,
        active: torch.Tensor | None = None,
        maxabs: torch.Tensor | None = None,
        ghost_rows: bool = False,
    ) -> torch.Tensor:
        batch, ld = self._state(uin)
        if L.rows_of(uout)[2] != ld or (u0 is not None and L.rows_of(u0)[2] != ld):
            raise ValueError("all stage arrays must share one row stride")
        dt_stride = 0 if (dt is None or dt.numel() == 1) else 1
        d = self.desc(batch, ld)
        L.check(
            "psk_ssprk33_stage",
            L.lib().psk_ssprk33_stage(
                ct.byref(d), stage, L.ptr(u0), L.ptr(uin), L.ptr(uout), L.ptr(dt), dt_stride,
                L.raw_ptr(active), L.ptr(self._lf_work(batch)), L.ptr(maxabs), int(ghost_rows), L.stream_ptr(),
            ),
        )
        return uout

    def rhs_axpby(self, u0: torch.Tensor, uin: torch.Tensor, dt: torch.Tensor, ca: float, cb: float, cc: float, *,
                  out: torch.Tensor | None = None, ghost_rows: bool = False) -> torch.Tensor:
        """``ca u0 + cb uin + cc dt L(uin)`` in one launch (``psk_rhs_axpby``): the stages of ForwardEuler / RK44 /
        CKRK45 (timestepping.py:289-405) with their combines fused into the right-hand side."""
        batch, ld = self._state(uin)
        res = _like(uin) if out is None else out
        if L.rows_of(u0)[2] != ld or L.rows_of(res)[2] != ld:
            raise ValueError("all stage arrays must share one row stride")
        d = self.desc(batch, ld)
        L.check(
            "psk_rhs_axpby",
            L.lib().psk_rhs_axpby(ct.byref(d), L.ptr(u0), L.ptr(uin), L.ptr(res), L.ptr(dt), 0 if dt.numel() == 1 else 1,
                                  float(ca), float(cb), float(cc), L.ptr(self._lf_work(batch)), int(ghost_rows),
                                  L.stream_ptr()),
        )
        return res

    def stage_lf(self, stage: int, u0: torch.Tensor, uin: torch.Tensor, uout: torch.Tensor, dt: torch.Tensor,
                 speed: torch.Tensor, maxabs: torch.Tensor) -> bool:
        """One fused stage with the global Lax-Friedrichs flux on periodic rows, the speed taken from the ``maxabs``
        output of the launch that produced ``uin`` instead of a reduction pass (``psk_ssprk33_stage_lf``);
        ``maxabs`` (zero-filled by the caller) receives the speed of the next stage.  ``False`` elsewhere."""
        batch, ld = self._state(uin)
        d = self.desc(batch, ld)
        rc = L.lib().psk_ssprk33_stage_lf(ct.byref(d), stage, L.ptr(u0), L.ptr(uin), L.ptr(uout), L.ptr(dt),
                                          0 if dt.numel() == 1 else 1, L.ptr(speed), L.ptr(maxabs), L.stream_ptr())
        if rc == L.E_UNSUPPORTED:
            return False
        L.check("psk_ssprk33_stage_lf", rc)
        return True

    def step_fused(
        self,
        u: torch.Tensor,
        uout: torch.Tensor,
        dt: torch.Tensor,
        *,
        active: torch.Tensor | None = None,
        maxabs: torch.Tensor | None = None,
        ghosts: Sequence[np.ndarray | torch.Tensor] | torch.Tensor | None = None,
    ) -> bool:
        """One whole SSPRK33 step (timestepping.py:312-320) in ONE launch, ``u -> uout`` (no aliasing),
        bit-identical to three :meth:`stage` calls.  Exists for the hot configurations only
        (``psk_ssprk33_step``; ``psk_ssprk33_step_bc`` for Dirichlet rows, ``ghosts`` = their data at
        ``t, t + dt, t + dt / 2``, default the data of :meth:`set_ghost` at all three): returns ``False`` --
        nothing launched -- anywhere else, and the caller runs the three stages.  Rows with ``active == 0`` are
        copied to ``uout``."""
        batch, ld = self._state(u)
        if L.rows_of(uout)[2] != ld:
            raise ValueError("all stage arrays must share one row stride")
        d = self.desc(batch, ld)
        if self.bc in ("dirichlet", "neumann"):
            g3 = self.ghost3(ghosts)
            if g3 is None:
                raise ValueError("Dirichlet / Neumann rows need boundary data (set_ghost or ghosts=)")
            if g3.dim() == 3 and g3.shape[1] != batch:
                raise ValueError("per-row ghost data does not match the batch size")
            d.ghost, d.ghost_ld = L.ptr(g3), (0 if g3.dim() == 2 else 2 * self.g)
            rc = L.lib().psk_ssprk33_step_bc(
                ct.byref(d), L.ptr(u), L.ptr(uout), L.ptr(dt), 0 if dt.numel() == 1 else 1, L.ptr(g3),
                L.raw_ptr(active), L.ptr(maxabs), None, None, L.stream_ptr(),
            )
            if rc == L.E_UNSUPPORTED:
                return False
            L.check("psk_ssprk33_step_bc", rc)
            return True
        if self.equation != "burgers" and not self._vel_periodic:
            return False
        rc = L.lib().psk_ssprk33_step(
            ct.byref(d), L.ptr(u), L.ptr(uout), L.ptr(dt), 0 if dt.numel() == 1 else 1,
            L.raw_ptr(active), L.ptr(maxabs), L.stream_ptr(),
        )
        if rc == L.E_UNSUPPORTED:
            return False
        L.check("psk_ssprk33_step", rc)
        return True

    _ENDS_W = 32  # cells kept next to each row end by ssprk33_advance: >= 3 stages x (g + 3) cells of dependence

    def _ends_path(self) -> "HotPath":
        """this path on the ``2 x _ENDS_W`` cells next to the row ends (the cut in the middle is never looked at)"""
        if self._ends is None:
            W, g, nx = self._ENDS_W, self.g, self.nx
            cut = lambda a: None if a is None else torch.cat((a[: g + W], a[nx - g - W :]))  # noqa: E731
            self._ends = HotPath(equation=self.equation, flux=self.flux, rec=self.rec, bc=self.bc, n=2 * W, g=g,
                                 dx=self.dx, eps=self.eps, math=self.math, nu=cut(self._nu), velocity=cut(self._vel),
                                 device=self.device, delta=self.delta)
        return self._ends

    def ssprk33_advance(self, u: torch.Tensor, dt: torch.Tensor, *,
                        ghosts: Sequence[np.ndarray | torch.Tensor] | None = None) -> torch.Tensor:
        """``advance(SSPRK33)`` of the reference (timestepping.py:312-320) INCLUDING the by-products it leaves in the
        ghost cells of the result: the whole-step kernel writes the interior in one launch, and the three stage
        launches run only on the ``2 x 32`` cells next to the row ends (the ghost cells of the result depend on
        nothing farther than ``3 (g + 3)`` cells away), whose ghost cells are copied over.  Falls back to the three
        full stage launches wherever the whole-step kernel does not exist."""
        W, g, nx = self._ENDS_W, self.g, self.nx
        # (global Lax-Friedrichs: the speed is a maximum over the whole row, the row ends alone do not know it)
        if self.math != "fast" or self.n < 4 * W or self.flux == "lf":
            return self.ssprk33_step(u, dt, ghosts=ghosts, ghost_rows=True)
        # the whole-step kernel wants 16-byte aligned interiors: a plain (batch, nx) array of the caller is copied
        # once into the padded row layout; the result is returned as a view of that layout, so the next advance of
        # a step() loop finds its input aligned and copies nothing
        ua = self.aligned(u)
        out = self._aligned_like(u)
        if not self.step_fused(ua, out, dt, ghosts=ghosts):
            return self.ssprk33_step(u, dt, ghosts=ghosts, ghost_rows=True)
        ends = self._ends_path()
        if ghosts is None and self._ghost is not None:
            ends._ghost, ends._ghost3, ends._ghost_ld = self._ghost, None, self._ghost_ld
        r = ends.ssprk33_step(torch.cat((ua[..., : g + W], ua[..., nx - g - W :]), dim=-1), dt, ghosts=ghosts,
                              ghost_rows=True)
        out[..., :g] = r[..., :g]
        out[..., nx - g :] = r[..., 2 * W + g :]
        return out

    def _rows_aligned(self, u: torch.Tensor) -> bool:
        """whether ``u`` already sits in the padded row layout (the arrays of _aligned_like and of EnsembleSolver)"""
        from .ensemble import row_layout

        return (u.data_ptr() + 8 * self.g) % 16 == 0 and (u.dim() == 1 or L.rows_of(u)[2] == row_layout(self.n, self.g)[1])

    def aligned(self, u: torch.Tensor) -> torch.Tensor:
        """``u`` itself if it sits in the padded row layout, else a copy that does"""
        if self._rows_aligned(u):
            return u
        ua = self._aligned_like(u)
        ua.copy_(u)
        return ua

    def _aligned_like(self, u: torch.Tensor) -> torch.Tensor:
        """uninitialised ``u.shape`` view of storage in the padded row layout (ensemble.row_layout)"""
        from .ensemble import row_layout

        col0, ld = row_layout(self.n, self.g)
        store = torch.empty(tuple(u.shape[:-1]) + (ld,), dtype=torch.float64, device=u.device)
        return store[..., col0 : col0 + self.nx]

    def step_fused_stages(self, u: torch.Tensor, k1: torch.Tensor, k2: torch.Tensor, uout: torch.Tensor | None,
                          dt: torch.Tensor) -> bool:
        """The stage values ``k1, k2`` of the SSPRK33 step from ``u`` (timestepping.py:314-317) and, if
        ``uout`` is given, the new state, in ONE launch (``psk_ssprk33_step_stages``): the recomputation
        of the reverse sweep.  ``False`` -- nothing launched -- outside its configuration."""
        batch, ld = self._state(u)
        for a in (k1, k2, uout):
            if a is not None and L.rows_of(a)[2] != ld:
                raise ValueError("all stage arrays must share one row stride")
        d = self.desc(batch, ld)
        if self.bc in ("dirichlet", "neumann"):  # rows with boundary data (the data of set_ghost at all three stage times)
            g3 = self.ghost3()
            if g3 is None or (g3.dim() == 3 and g3.shape[1] != batch):
                return False
            d.ghost, d.ghost_ld = L.ptr(g3), (0 if g3.dim() == 2 else 2 * self.g)
            rc = L.lib().psk_ssprk33_step_bc(ct.byref(d), L.ptr(u), L.ptr(uout), L.ptr(dt), 0 if dt.numel() == 1 else 1,
                                             L.ptr(g3), None, None, L.ptr(k1), L.ptr(k2), L.stream_ptr())
            if rc == L.E_UNSUPPORTED:
                return False
            L.check("psk_ssprk33_step_bc", rc)
            return True
        rc = L.lib().psk_ssprk33_step_stages(
            ct.byref(d), L.ptr(u), L.ptr(k1), L.ptr(k2), L.ptr(uout), L.ptr(dt), 0 if dt.numel() == 1 else 1,
            L.stream_ptr(),
        )
        if rc == L.E_UNSUPPORTED:
            return False
        L.check("psk_ssprk33_step_stages", rc)
        return True

    def reverse_step_supported(self) -> bool:
        """whether :meth:`reverse_step_fused` exists for this scheme (the conditions of psk_ssprk33_step_adjoint /
        psk_ssprk33_step_adjoint_bc; the arrays must also be 16-byte aligned with an even row stride, which
        EnsembleSolver guarantees)"""
        return (self.equation == "burgers" and self.flux == "rusanov" and self.rec == "wenojs53" and self.math == "fast"
                and ((self.bc == "periodic" and self.g >= 3) or (self.bc == "none" and self.g >= 16)
                     or (self.bc == "dirichlet" and self.g == 3))
                and self._nu is None and self.n % 2 == 0 and self.n >= 8)

    def reverse_step_fused(self, u: torch.Tensor, p: torch.Tensor, dt: torch.Tensor, out: torch.Tensor, *,
                           stages: tuple[torch.Tensor, torch.Tensor] | None = None,
                           ghosts: Sequence[np.ndarray | torch.Tensor] | torch.Tensor | None = None) -> bool:
        """``out = (d advance(dt, u) / d u)^T p`` for one SSPRK33 step from the checkpointed state ``u`` in ONE
        launch (``psk_ssprk33_step_adjoint``: ``k1, k2`` recomputed and the three adjoint stages applied inside
        the kernel; timestepping.py:198-209 without the dense Jacobian).  Interior cells only (periodic rings;
        Dirichlet rows -- ``psk_ssprk33_step_adjoint_bc``, ``ghosts`` as in :meth:`step_fused` -- whose boundary
        data carry no cotangent: the ghost cells of ``p`` count as zero); ``stages``: optional arrays that receive the
        recomputed ``k1, k2``.  ``False`` -- nothing launched -- outside its configuration (the caller then runs
        :meth:`ssprk33_step_adjoint`)."""
        batch, ld = self._state(u)
        for a in (p, out) + (tuple(stages) if stages is not None else ()):
            if L.rows_of(a)[2] != ld:
                raise ValueError("all arrays must share one row stride")
        d = self.desc(batch, ld)
        k1, k2 = stages if stages is not None else (None, None)
        if self.bc == "dirichlet":
            g3 = self.ghost3(ghosts)
            if g3 is None:
                raise ValueError("Dirichlet rows need boundary data (set_ghost or ghosts=)")
            if g3.dim() == 3 and g3.shape[1] != batch:
                raise ValueError("per-row ghost data does not match the batch size")
            d.ghost, d.ghost_ld = L.ptr(g3), (0 if g3.dim() == 2 else 2 * self.g)
            rc = L.lib().psk_ssprk33_step_adjoint_bc(
                ct.byref(d), L.ptr(u), L.ptr(p), L.ptr(dt), 0 if dt.numel() == 1 else 1, L.ptr(g3), L.ptr(out),
                L.ptr(k1), L.ptr(k2), L.stream_ptr(),
            )
            if rc == L.E_UNSUPPORTED:
                return False
            L.check("psk_ssprk33_step_adjoint_bc", rc)
            return True
        rc = L.lib().psk_ssprk33_step_adjoint(
            ct.byref(d), L.ptr(u), L.ptr(p), L.ptr(dt), 0 if dt.numel() == 1 else 1, L.ptr(out), L.ptr(k1), L.ptr(k2),
            L.stream_ptr(),
        )
        if rc == L.E_UNSUPPORTED:
            return False
        L.check("psk_ssprk33_step_adjoint", rc)
        return True

    def ssprk33_step(
        self,
        u: torch.Tensor,
        dt: torch.Tensor,
        *,
        out: torch.Tensor | None = None,
        ghosts: Sequence[np.ndarray | torch.Tensor] | None = None,
        active: torch.Tensor | None = None,
        maxabs: torch.Tensor | None = None,
        ghost_rows: bool = False,
        keep_stages: bool = False,
    ) -> torch.Tensor | tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One SSPRK33 step (timestepping.py:312-320) as three fused launches.

        ``ghosts``: Dirichlet / Neumann data at the stage times ``t, t + dt, t + dt/2``.
        """
        k1, k2 = _like(u), _like(u)
        res = _like(u) if out is None else out
        if active is not None:
            res.copy_(u)  # rows that are switched off keep their state
        for s, (src, dst) in enumerate(((u, k1), (k1, k2), (k2, res)), start=1):
            if ghosts is not None:
                self.set_ghost(ghosts[s - 1])
            self.stage(s, u, src, dst, dt, active=active, maxabs=maxabs if s == 3 else None, ghost_rows=ghost_rows)
        return (k1, k2, res) if keep_stages else res

    # }}}
