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

from dataclasses import dataclass
from functools import singledispatch
from typing import Any, Callable, ClassVar, Iterator

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
    return hp.ssprk33_step(u, _as_dt(dt, u), ghosts=ghosts, ghost_rows=True)


# {{{ RK44 / CKRK45 (timestepping.py:325-405): generic steppers -- each RHS is one fused
# apply_operator launch (boundary fill + WENO + flux + flux difference), the stage combines are
# the reference's own array expressions


@dataclass(frozen=True)
class RK44(Stepper):
    """The classic fourth-order Runge-Kutta method with 4 stages (timestepping.py:328-343)."""


@advance.register(RK44)
def _advance_rk44(stepper: RK44, dt: ScalarLike, t: ScalarLike, u: Array) -> Array:
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
        from .schemes import predict_timestep

        dt = theta * float(predict_timestep(scheme, grid, bc, 0.0, u))
        nsteps, dt = predict_maxit_from_timestep(tfinal, dt)
        out = hp.solve_rows(u, fixed_dt=dt, max_steps=nsteps, record_dt=True, tape=checkpoint)
    steps = out["steps"]
    if bool((steps < 0).any()):
        raise ValueError("Time step is not finite.")  # timestepping.py:144-145
    nmax = int(steps.max())
    return {"u": u, "t": out["t"], "iteration": steps, "dt": out["dt"][..., :nmax],
            "states": None if out["tape"] is None else out["tape"][: nmax + 1]}
