"""Linear advection, non-conservative form (``pyshocks/advection/schemes.py:27-129``)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import torch

from ..binding import hotpath_for, kernel_spec
from ..schemes import (
    Boundary,
    FiniteVolumeSchemeBase,
    SchemeBase,
    apply_operator,
    numerical_flux,
    predict_timestep,
)

Array = torch.Tensor
ScalarLike = Any


@dataclass(frozen=True, eq=False)
class AdvectionScheme(SchemeBase):
    velocity: Array | None
    """Advection velocity at the cell centres (ghosts included)."""


@predict_timestep.register(AdvectionScheme)
def _predict_timestep_advection(scheme: AdvectionScheme, grid: Any, bc: Boundary, t: ScalarLike, u: Array) -> Array:
    assert scheme.velocity is not None
    # advection/schemes.py:51-59: state independent
    amax = hotpath_for(scheme, grid, bc).max_abs(scheme.velocity, 1)[0]
    return grid.dx_min / amax


@apply_operator.register(AdvectionScheme)
def _apply_operator_advection(scheme: AdvectionScheme, grid: Any, bc: Boundary, t: ScalarLike, u: Array) -> Array:
    assert scheme.velocity is not None
    # advection/schemes.py:62-73: -velocity * (f[1:] - f[:-1]) / dx, one fused launch
    from ..timestepping import _trace_apply_operator

    out = hotpath_for(scheme, grid, bc, t).apply_operator(u)
    _trace_apply_operator(scheme, grid, bc, t, u, out)
    return out


@dataclass(frozen=True, eq=False)
class FiniteVolumeScheme(AdvectionScheme, FiniteVolumeSchemeBase):
    pass


@dataclass(frozen=True, eq=False)
class Godunov(FiniteVolumeScheme):
    """Upwind scheme (advection/schemes.py:100-129)."""


Upwind = Godunov


@kernel_spec.register(Godunov)
def _spec(scheme: Godunov) -> dict:
    assert scheme.velocity is not None
    return {"equation": "advection", "flux": "godunov", "alpha": 1.0, "velocity": scheme.velocity}


def upwind_flux(scheme: AdvectionScheme, grid: Any, bc: Boundary, u: Array) -> Array:
    assert scheme.velocity is not None
    assert scheme.rec is not None
    assert u.shape[-1] == grid.x.shape[0]
    from ..binding import NoBoundary

    return hotpath_for(scheme, grid, NoBoundary()).numerical_flux(u)


@numerical_flux.register(Godunov)
def _numerical_flux_advection_godunov(scheme: Godunov, grid: Any, bc: Boundary, t: ScalarLike, u: Array) -> Array:
    return upwind_flux(scheme, grid, bc, u)
