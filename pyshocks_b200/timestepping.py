"""Time stepping (``pyshocks/timestepping.py``): ``Stepper`` / ``SSPRK33`` / ``ForwardEuler``,
the ``step`` generator (:89-152), ``adjoint_step`` (:155-215), ``advance`` (:218-233, :297-320)
and the fixed-step helpers (:242-280).

``advance(SSPRK33)`` is three fused stage launches (boundary fill + WENO + flux + flux
difference + stage combine each) when the stepper's ``source`` was made from
``apply_operator`` with :func:`pyshocks_b200.jit` or :func:`bind_operator`; otherwise the
generic path of the reference (three RHS evaluations plus axpys) is used.

``adjoint_step`` replaces the dense ``jax.jacfwd`` Jacobian of one step (O(nx^2), :174, :205-206)
with the matrix-free transposed-stencil kernels: recompute ``k1, k2`` from the checkpointed
state, then three fused adjoint stage launches.
"""

from __future__ import annotations

import functools
from dataclasses import dataclass
from functools import singledispatch
from typing import Any, Callable, ClassVar, Iterator

import numpy as np
import torch

from .checkpointing import Checkpoint, load, save

Array = torch.Tensor
ScalarLike = Any


# {{{ interface


@dataclass(frozen=True)
class StepCompleted:
    t: Any
    tfinal: Any
    dt: Any
    iteration: int
    u: Array

    def __str__(self) -> str:
        return f"[{self.iteration:5d}] t = {float(self.t):.5e} / {float(self.tfinal):.5e} dt {float(self.dt):.5e}"


@dataclass(frozen=True)
class AdjointStepCompleted(StepCompleted):
    p: Array


@dataclass(frozen=True)
class Stepper:
    predict_timestep: Callable[[ScalarLike, Array], Any]
    source: Callable[[ScalarLike, Array], Array]
    checkpoint: Checkpoint | None


class BoundOperator:
    """``(t, u) -> apply_operator(scheme, grid, bc, t, u)`` that remembers what it is bound to,
    so that ``advance`` / ``adjoint_step`` can launch the fused kernels."""

    def __init__(self, scheme: Any, grid: Any, bc: Any) -> None:
        self.scheme, self.grid, self.bc = scheme, grid, bc

    def __call__(self, t: ScalarLike, u: Array) -> Array:
        from .schemes import apply_operator

        return apply_operator(self.scheme, self.grid, self.bc, t, u)


def bind_operator(scheme: Any, grid: Any, bc: Any) -> BoundOperator:
    return BoundOperator(scheme, grid, bc)


_TRACE: list | None = None


def _trace_apply_operator(scheme: Any, grid: Any, bc: Any, t: Any, u: Any, out: Any) -> None:
    if _TRACE is not None:
        _TRACE.append((scheme, grid, bc, t, u, out))


def jit(fun: Callable | None = None, **kwargs: Any) -> Callable:
    """Stand-in for ``jax.jit`` in user code such as ``source=jax.jit(_apply_operator)``
    (examples/burgers.py:173-176).  The first call records what the function does: if it is
    exactly one ``apply_operator(scheme, grid, bc, t, u)`` on its own arguments, the wrapper
    exposes the bound triple and ``advance`` fuses the stage; any other function just runs."""
    if fun is None:
        return lambda f: jit(f, **kwargs)

    class _Jitted:
        def __init__(self) -> None:
            self._bound: BoundOperator | None = None
            self._traced = False
            self.__wrapped__ = fun

        @property
        def bound(self) -> BoundOperator | None:
            return self._bound

        def __call__(self, *args: Any, **kw: Any) -> Any:
            global _TRACE
            if self._traced:
                return fun(*args, **kw)
            self._traced = True
            prev, _TRACE = _TRACE, []
            try:
                out = fun(*args, **kw)
                calls = _TRACE
            finally:
                _TRACE = prev
            if len(calls) == 1 and len(args) == 2 and not kw:
                scheme, grid, bc, t, u, res = calls[0]
                if res is out and u is args[1] and (t is args[0] or t == args[0]):
                    self._bound = BoundOperator(scheme, grid, bc)
            return out

    return _Jitted()


def _bound_of(source: Any, probe: tuple | None = None) -> BoundOperator | None:
    if isinstance(source, BoundOperator):
        return source
    if isinstance(source, functools.partial) and not source.keywords and len(source.args) == 3:
        from .schemes import apply_operator

        if source.func is apply_operator:  # partial(apply_operator, scheme, grid, bc)
            return BoundOperator(*source.args)
    b = getattr(source, "bound", None)
    if b is None and probe is not None and hasattr(source, "_traced") and not source._traced:
        source(*probe)  # first call of a jit() wrapper: trace it
        b = source.bound
    return b


def _as_dt(dt: Any, like: Array) -> Array:
    if isinstance(dt, torch.Tensor):
        return dt.reshape(-1).to(dtype=torch.float64, device=like.device)
    return torch.full((1,), float(dt), dtype=torch.float64, device=like.device)


def step(
    stepper: Stepper,
    u0: Array,
    *,
    maxit: int | None = None,
    tstart: ScalarLike = 0.0,
    tfinal: ScalarLike | None = None,
) -> Iterator[StepCompleted]:
    """timestepping.py:89-152.  Like the reference, the host reads ``dt`` every step."""
    if tfinal is None:
        tfinal = float(torch.finfo(u0.dtype).max)
    m = 0
    t = torch.tensor(float(tstart), dtype=u0.dtype, device=u0.device)
    tfinal = torch.tensor(float(tfinal), dtype=u0.dtype, device=u0.device)
    u = u0
    yield StepCompleted(t=t, tfinal=tfinal, dt=torch.zeros((), dtype=u0.dtype, device=u0.device), iteration=m, u=u)
    while True:
        if stepper.checkpoint is not None:
            save(stepper.checkpoint, m, {"m": m, "t": t, "u": u})
        if tfinal is not None and t >= tfinal:
            break
        if maxit is not None and m >= maxit:
            break
        dt = stepper.predict_timestep(t, u)
        if not isinstance(dt, torch.Tensor):
            dt = torch.tensor(float(dt), dtype=u0.dtype, device=u0.device)
        dt = dt.reshape(())
        if tfinal != float("inf"):
            dt_min = tfinal - t
            dt = (dt if dt < dt_min else dt_min) + 1.0e-15
        if not torch.isfinite(dt):
            raise ValueError(f"Time step is not finite: {dt!r}.")
        u = advance(stepper, dt, t, u)
        m += 1
        t = t + dt
        yield StepCompleted(t=t, tfinal=tfinal, dt=dt, iteration=m, u=u)


def adjoint_step(
    stepper: Stepper,
    p0: Array,
    *,
    maxit: int,
    apply_boundary: Callable[[Any, Array, Array], Array] | None = None,
) -> Iterator[AdjointStepCompleted]:
    """timestepping.py:155-215 with ``jac.T @ p`` evaluated matrix-free on the GPU."""
    if stepper.checkpoint is None:
        raise ValueError("Adjoint time stepping requires a checkpoint.")
    if not isinstance(stepper, SSPRK33):
        raise NotImplementedError(f"adjoint of {type(stepper).__name__} (only SSPRK33 is on the hot path)")
    chk = load(stepper.checkpoint, maxit)
    assert chk["m"] == maxit
    bound = _bound_of(stepper.source, (chk["t"], chk["u"]))
    if bound is None:
        raise NotImplementedError(
            "adjoint_step needs a source built from apply_operator: "
            "use pyshocks_b200.jit(lambda t, u: apply_operator(scheme, grid, bc, t, u)) or bind_operator(...)"
        )
    from .binding import ghost_data, hotpath_for

    t = tfinal = chk["t"]
    p = p0
    if apply_boundary is not None:
        p = apply_boundary(chk["t"], chk["u"], p)
    yield AdjointStepCompleted(
        t=t, tfinal=tfinal, dt=torch.zeros((), dtype=p.dtype, device=p.device), iteration=maxit, u=chk["u"], p=p
    )
    for m in range(maxit - 1, -1, -1):
        chk = load(stepper.checkpoint, m)
        dt = t - chk["t"]
        assert chk["m"] == m
        hp = hotpath_for(bound.scheme, bound.grid, bound.bc)
        tm = chk["t"]
        ghosts = None
        if ghost_data(bound.bc, bound.grid, tm) is not None:
            ghosts = [ghost_data(bound.bc, bound.grid, tt) for tt in (tm, tm + dt, tm + 0.5 * dt)]
        p = hp.ssprk33_step_adjoint(chk["u"], _as_dt(dt, p), p, ghosts=ghosts)
        if apply_boundary is not None:
            p = apply_boundary(chk["t"], chk["u"], p)
        t = chk["t"]
        yield AdjointStepCompleted(t=t, tfinal=tfinal, dt=dt, iteration=m, u=chk["u"], p=p)


@singledispatch
def advance(stepper: Stepper, dt: ScalarLike, t: ScalarLike, u: Array) -> Array:
    """timestepping.py:218-233."""
    raise NotImplementedError(type(stepper).__name__)


# }}}

# {{{ fixed time step helpers (timestepping.py:242-280)


def predict_timestep_from_maxit(tfinal: float, maxit: int) -> tuple[int, float]:
    return maxit, tfinal / maxit + 1.0e-15


def predict_maxit_from_timestep(tfinal: float, dt: float) -> tuple[int, float]:
    maxit = int(tfinal / float(dt))
    return maxit, tfinal / maxit + 1.0e-15


def predict_timestep_from_resolutions(a: float, b: float, resolutions: list[int], *, umax: float = 1.0, p: int = 1) -> float:
    dx = (b - a) / max(resolutions)
    return dx**p / umax


# }}}


@dataclass(frozen=True)
class ForwardEuler(Stepper):
    """Forward Euler (timestepping.py:289-301): falls out of the fused stage-1 kernel."""


@advance.register(ForwardEuler)
def _advance_forward_euler(stepper: ForwardEuler, dt: ScalarLike, t: ScalarLike, u: Array) -> Array:
    bound = _bound_of(stepper.source, (t, u))
    if bound is None:
        return u + dt * stepper.source(t, u)
    from .binding import hotpath_for
    from .path import _like

    hp = hotpath_for(bound.scheme, bound.grid, bound.bc, t)
    if hp.math == "fast":  # arrays in the padded row layout: vector loads; nothing to copy from the 2nd step on
        u = hp.aligned(u)
        return hp.stage(1, u, u, hp._aligned_like(u), _as_dt(dt, u), ghost_rows=True)
    return hp.stage(1, u, u, _like(u), _as_dt(dt, u), ghost_rows=True)


@dataclass(frozen=True)
class SSPRK33(Stepper):
    """The optimal third-order SSP Runge-Kutta method with three stages (timestepping.py:307-320)."""


@advance.register(SSPRK33)
def _advance_ssprk33(stepper: SSPRK33, dt: ScalarLike, t: ScalarLike, u: Array) -> Array:
    bound = _bound_of(stepper.source, (t, u))
    if bound is None:
        fn = stepper.source
        k1 = u + dt * fn(t, u)
        k2 = 3.0 / 4.0 * u + 1.0 / 4.0 * (k1 + dt * fn(t + dt, k1))
        return 1.0 / 3.0 * u + 2.0 / 3.0 * (k2 + dt * fn(t + 0.5 * dt, k2))
    from .binding import ghost_data, hotpath_for

    hp = hotpath_for(bound.scheme, bound.grid, bound.bc)
    ghosts = None
    if ghost_data(bound.bc, bound.grid, t) is not None:
        ghosts = [ghost_data(bound.bc, bound.grid, tt) for tt in (t, t + dt, t + 0.5 * dt)]
    return hp.ssprk33_advance(u, _as_dt(dt, u), ghosts=ghosts)


# {{{ RK44 / CKRK45 (timestepping.py:325-405): generic steppers -- each RHS is one fused
# apply_operator launch (boundary fill + WENO + flux + flux difference), the stage combines are
# the reference's own array expressions


@dataclass(frozen=True)
class RK44(Stepper):
    """The classic fourth-order Runge-Kutta method with 4 stages (timestepping.py:328-343)."""


def _fused_binding(stepper: Stepper, t: ScalarLike, u: Array):
    """the (scheme, grid, bc) of a source that is exactly one apply_operator call, in FAST math: the stage combines
    of RK44 / CKRK45 are then fused into the right-hand side kernel (psk_rhs_axpby).  STRICT math keeps the
    reference's own array expressions below -- bit-identical to its advance, which a fused combine cannot be."""
    from . import config

    if config.MATH != "fast":
        return None
    return _bound_of(stepper.source, (t, u))


@advance.register(RK44)
def _advance_rk44(stepper: RK44, dt: ScalarLike, t: ScalarLike, u: Array) -> Array:
    bound = _fused_binding(stepper, t, u)
    if bound is not None:
        # y2 = u + dt/2 L(t, u), y3 = u + dt/2 L(t + dt/2, y2), y4 = u + dt L(t + dt/2, y3),
        # u' = u + (k1 + 2 k2 + 2 k3 + k4) / 6 = (-u + y2 + 2 y3) / 3 + y4 / 3 + dt/6 L(t + dt, y4)
        from .binding import hotpath_for

        dtt = _as_dt(dt, u)
        hp = hotpath_for(bound.scheme, bound.grid, bound.bc, t)
        u = hp.aligned(u)  # (arrays in the padded row layout: vector loads in the kernels; no copy from the 2nd step on)
        y2 = hp.rhs_axpby(u, u, dtt, 1.0, 0.0, 0.5, out=hp._aligned_like(u), ghost_rows=True)
        hp = hotpath_for(bound.scheme, bound.grid, bound.bc, t + dt / 2)
        y3 = hp.rhs_axpby(u, y2, dtt, 1.0, 0.0, 0.5, out=hp._aligned_like(u), ghost_rows=True)
        y4 = hp.rhs_axpby(u, y3, dtt, 1.0, 0.0, 1.0, out=hp._aligned_like(u), ghost_rows=True)
        acc = torch.add(y2, y3, alpha=2.0, out=y2).sub_(u)  # 3 x the part of u' that needs no further L
        hp = hotpath_for(bound.scheme, bound.grid, bound.bc, t + dt)
        return hp.rhs_axpby(acc, y4, dtt, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0, out=y3, ghost_rows=True)
    fn = stepper.source
    k1 = dt * fn(t, u)
    k2 = dt * fn(t + dt / 2, u + k1 / 2)
    k3 = dt * fn(t + dt / 2, u + k2 / 2)
    k4 = dt * fn(t + dt, u + k3)
    # a true division: torch divides a CUDA tensor by a Python scalar as a multiplication by its reciprocal,
    # which is not the reference's `/ 6` in the last bit
    six = torch.full((), 6.0, dtype=u.dtype, device=u.device)
    return u + torch.div(k1 + 2 * k2 + 2 * k3 + k4, six)


@dataclass(frozen=True)
class CKRK45(Stepper):
    """Low-storage five-stage fourth-order method of Carpenter and Kennedy (timestepping.py:352-405)."""

    a: ClassVar[tuple[float, ...]] = (
        0.0,
        -567301805773 / 1357537059087,
        -2404267990393 / 2016746695238,
        -3550918686646 / 2091501179385,
        -1275806237668 / 842570457699,
    )
    b: ClassVar[tuple[float, ...]] = (
        1432997174477 / 9575080441755,
        5161836677717 / 13612068292357,
        1720146321549 / 2090206949498,
        3134564353537 / 4481467310338,
        2277821191437 / 14882151754819,
    )
    c: ClassVar[tuple[float, ...]] = (
        0.0,
        1432997174477 / 9575080441755,
        2526269341429 / 6820363962896,
        2006345519317 / 3224310063776,
        2802321613138 / 2924317926251,
    )


@advance.register(CKRK45)
def _advance_ckrk45(stepper: CKRK45, dt: ScalarLike, t: ScalarLike, u: Array) -> Array:
    bound = _fused_binding(stepper, t, u)
    if bound is not None:
        # k <- a_i k + dt L(t + c_i dt, p) in one launch, p <- p + b_i k
        from .binding import hotpath_for

        dtt = _as_dt(dt, u)
        hp = hotpath_for(bound.scheme, bound.grid, bound.bc, t)
        p = k = hp.aligned(u)
        for i in range(len(stepper.a)):
            hp = hotpath_for(bound.scheme, bound.grid, bound.bc, t + stepper.c[i] * dt)
            # (k is both u0 and the output from the 2nd stage on: every cell reads its own u0 only)
            k = hp.rhs_axpby(k, p, dtt, stepper.a[i], 0.0, 1.0, out=hp._aligned_like(u) if i == 0 else k, ghost_rows=True)
            p = torch.add(p, k, alpha=stepper.b[i], out=hp._aligned_like(u) if i == 0 else p)
        return p
    fn = stepper.source
    p = k = u
    for i in range(len(stepper.a)):
        k = stepper.a[i] * k + dt * fn(t + stepper.c[i] * dt, p)
        p = p + stepper.b[i] * k
    return p


# }}}


def solve(
    scheme: Any,
    grid: Any,
    bc: Any,
    u0: Array,
    *,
    tfinal: float,
    theta: float = 1.0,
    maxit: int = 1 << 14,
    checkpoint: bool = False,
) -> dict:
    """``for event in step(SSPRK33(...), u0, tfinal=tfinal): pass`` as ONE kernel launch
    (``psk_solve_rows``) for grids that fit in shared memory and boundary data that do not depend on
    time: the CFL reduction, the dt clamp and the three fused stages of every step run on the
    device.  Returns ``{"u", "t", "iteration", "dt", "states"}``; ``dt`` is the per-step history and
    ``states`` (with ``checkpoint=True``) what an ``InMemoryCheckpoint`` would hold.

    Not in the reference (its loop is a Python generator with a host read of dt per step,
    timestepping.py:139-150); results are identical to the step-by-step path in STRICT mode."""
    from .binding import hotpath_for
    from .burgers.schemes import BurgersScheme, Rusanov

    hp = hotpath_for(scheme, grid, bc, 0.0)
    u = u0.clone()
    if isinstance(scheme, BurgersScheme):
        alpha = scheme.alpha if isinstance(scheme, Rusanov) else 1.0
        cfl_scale = 0.5 * grid.h ** (2 - alpha) if isinstance(scheme, Rusanov) else 0.5 * grid.h
        out = hp.solve_rows(u, tfinal=tfinal, theta=theta, cfl_scale=cfl_scale, max_steps=maxit,
                            record_dt=True, tape=checkpoint)
    else:
        # state-independent time step (advection/schemes.py:51-59): the dt sequence of step() -- clamped at tfinal,
        # + 1e-15 (timestepping.py:139-142) -- and the boundary data at every stage time are known in advance
        from .schemes import predict_timestep

        dts, ts = step_sizes(theta * float(predict_timestep(scheme, grid, bc, 0.0, u)), tfinal, maxit)
        dts_dev = torch.tensor(dts, dtype=torch.float64, device=u.device)
        table = ghost_table(bc, grid, ts, dts)
        tape = hp.steps_tape(u, dts_dev, table) if checkpoint else None
        if tape is not None:
            # one whole-step launch per step, enqueued from one call, every state written straight onto the tape.
            # The whole-step kernel writes INTERIOR cells only: the ghost cells of these states are zero, not the
            # values the reference's full-array advance leaves there (zero-padded stencils plus the ghost values
            # carried over from step to step, schemes.py:346) -- numbers no later step and no adjoint step reads:
            # every right-hand side starts by overwriting the ghost cells (schemes.py:343)
            u = tape[len(dts), 0].clone() if u.dim() == 1 else tape[len(dts)].clone()
            out = {"t": torch.full((tape.shape[1],), ts[-1] + dts[-1], dtype=torch.float64, device=u.device),
                   "steps": torch.full((tape.shape[1],), len(dts), dtype=torch.int32, device=u.device), "tape": tape}
        else:
            out = hp.solve_rows_tables(u, dts_dev, table, tape=checkpoint)
        out["dt"] = dts_dev[None, :]
        out["ghost_table"], out["ts"] = table, ts
    steps = out["steps"]
    if bool((steps < 0).any()):
        raise ValueError("Time step is not finite.")  # timestepping.py:144-145
    nmax = int(steps.max())
    tape = out["tape"]
    if tape is not None and tape.shape[-1] != hp.nx:
        tape = tape[..., : hp.nx]
    res = {"u": u, "t": out["t"], "iteration": steps, "dt": out["dt"][..., :nmax],
           "states": None if tape is None else tape[: nmax + 1]}
    if "ghost_table" in out:
        res["ghost_table"], res["ts"] = out["ghost_table"], out["ts"]
    return res


def step_sizes(dt_cfl: float, tfinal: float, maxit: int | None = None) -> tuple[list[float], list[float]]:
    """The ``dt`` and ``t`` sequences of :func:`step` for a state-independent ``predict_timestep``
    (timestepping.py:128-150: ``dt = min(dt, tfinal - t) + 1e-15`` until ``t >= tfinal`` or ``maxit`` steps)."""
    if not np.isfinite(dt_cfl):
        raise ValueError(f"Time step is not finite: {dt_cfl!r}.")
    dts: list[float] = []
    ts: list[float] = []
    t = 0.0
    while not t >= tfinal and (maxit is None or len(dts) < maxit):
        dt_min = tfinal - t
        dt = (dt_cfl if dt_cfl < dt_min else dt_min) + 1.0e-15
        ts.append(t)
        dts.append(dt)
        t = t + dt
    return dts, ts


def ghost_table(bc: Any, grid: Any, ts: list[float], dts: list[float]) -> torch.Tensor | None:
    """``(nsteps, 3, 2 g)`` boundary data at the stage times ``t, t + dt, t + dt / 2`` of every step
    (timestepping.py:314-319), evaluated with the user's ``g(t, x)`` exactly as ``apply_boundary`` does step
    by step (scalar.py:424-425, :490-498); ``None`` for boundaries without data (periodic)."""
    from .binding import ghost_data

    if not ts or ghost_data(bc, grid, ts[0]) is None:
        return None
    dev = grid.x.device
    times = [tt for t, dt in zip(ts, dts) for tt in (t, t + dt, t + 0.5 * dt)]

    def one(tt: float) -> torch.Tensor:
        gd = ghost_data(bc, grid, tt)
        return (gd if isinstance(gd, torch.Tensor) else torch.from_numpy(np.asarray(gd, dtype=np.float64))).to(dev)

    # one broadcast call per side where the user's g(t, x) is elementwise in (t, x) -- checked against the
    # time-by-time evaluation on a few rows -- instead of 3 x nsteps small calls
    from .scalar import DirichletBoundary, TwoSidedBoundary

    if isinstance(bc, TwoSidedBoundary) and isinstance(bc.left, DirichletBoundary) and isinstance(bc.right, DirichletBoundary):
        try:
            g, nx = grid.nghosts, grid.x.shape[0]
            T = torch.tensor(times, dtype=torch.float64, device=dev)[:, None]
            left = torch.as_tensor(bc.left.g(T, grid.x[None, :g]), dtype=torch.float64, device=dev)
            right = torch.as_tensor(bc.right.g(T, grid.x[None, nx - g :]), dtype=torch.float64, device=dev)
            if tuple(left.shape) == (len(times), g) and tuple(right.shape) == (len(times), g):
                table = torch.cat([left, right], dim=1)
                probe = sorted({0, 1, len(times) // 2, len(times) - 1})
                if all(torch.equal(table[k], one(times[k])) for k in probe):
                    return table.reshape(len(ts), 3, -1).contiguous()
        except Exception:  # noqa: BLE001  (a g(t, x) that does not broadcast: evaluate it time by time)
            pass
    return torch.stack([one(tt) for tt in times]).reshape(len(ts), 3, -1).contiguous()


def adjoint_solve(
    scheme: Any,
    grid: Any,
    bc: Any,
    forward: dict,
    p0: Array,
    *,
    p_boundary: Any = None,
    history: bool = False,
) -> dict:
    """``for event in adjoint_step(stepper, p0, maxit=..., apply_boundary=...): pass`` in ONE call
    (``psk_ssprk33_adjoint_sweep``) from the result of :func:`solve` with ``checkpoint=True``: every reverse
    step is six kernel launches enqueued back to back, no host round trip.  ``p_boundary``: the boundary
    condition the drivers impose on the adjoint variable after every step (homogeneous Dirichlet or
    Neumann; its data must not depend on time).  Returns ``{"p": p(0), "history": (steps, nx) or None}``.

    Not in the reference (its loop is a Python generator around a dense Jacobian, timestepping.py:155-215)."""
    from .binding import boundary_kind, ghost_data, hotpath_for
    from .path import HotPath

    hp = hotpath_for(scheme, grid, bc, 0.0)
    states = forward["states"]
    if states is None:
        raise ValueError("Adjoint time stepping requires a checkpoint.")  # timestepping.py:177-178
    tape = states if states.dim() == 3 else states[:, None, :]
    nsteps = tape.shape[0] - 1
    # like adjoint_step, every reverse step uses dt = t_{m+1} - t_m of the ACCUMULATED times (timestepping.py:200-202),
    # which is not always the forward dt_m in the last bit -- for the recomputed stages and the boundary data too
    fdts = [float(x) for x in forward["dt"].reshape(-1)[:nsteps].cpu().numpy()]
    ts = [0.0]
    for dt in fdts:
        ts.append(ts[-1] + dt)
    rdts = [ts[m + 1] - ts[m] for m in range(nsteps)]
    dts = torch.tensor(rdts, dtype=torch.float64, device=p0.device)
    table = forward.get("ghost_table")
    if table is not None and rdts != fdts:
        table = ghost_table(bc, grid, ts[:-1], rdts)
    p = p0.clone()
    pb = None
    if p_boundary is not None:
        pb = HotPath(equation="burgers", flux="rusanov", rec="constant", bc=boundary_kind(p_boundary), n=hp.n, g=hp.g,
                     dx=hp.dx, eps=0.0, device=p.device)
        gd = ghost_data(p_boundary, grid, 0.0)
        if gd is not None:
            pb.set_ghost(gd)
        p = pb.apply_boundary(p)  # timestepping.py:186-187
    # p gets the row layout of the tape (same stride, same position of the first cell in its row)
    batch = tape.shape[1]
    ld = tape.stride(1) if batch > 1 else max(tape.stride(1), hp.nx)
    col = tape.storage_offset() % ld if ld > hp.nx else 0
    if col + hp.nx > ld:
        col = 0
    pbuf = torch.zeros((batch, ld), dtype=torch.float64, device=p.device)
    pv = pbuf[:, col : col + hp.nx]
    pv.copy_(p if p.dim() == 2 else p[None, :])
    hist = hp.adjoint_sweep(tape, dts, pv, ghost_table=table, p_boundary=pb, history=history)
    out_p = pv[0].clone() if p0.dim() == 1 else pv.clone()
    return {"p": out_p, "history": None if hist is None else (hist[:, 0] if p0.dim() == 1 else hist)}
