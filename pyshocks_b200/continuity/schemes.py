"""Continuity equation, conservative form (``pyshocks/continuity/schemes.py:27-110``)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import torch

from ..binding import hotpath_for, kernel_spec
from ..schemes import Boundary, ConservationLawScheme, SchemeBase, flux, numerical_flux, predict_timestep

Array = torch.Tensor
ScalarLike = Any


@dataclass(frozen=True, eq=False)
class ContinuityScheme(SchemeBase):
    velocity: Array | None


@flux.register(ContinuityScheme)
def _flux_continuity(scheme: ContinuityScheme, t: ScalarLike, x: Array, u: Array) -> Array:
    assert scheme.velocity is not None
    return scheme.velocity * u  # continuity/schemes.py:39-45


@predict_timestep.register(ContinuityScheme)
def _predict_timestep_continuity(scheme: ContinuityScheme, grid: Any, bc: Boundary, t: ScalarLike, u: Array) -> Array:
    assert scheme.velocity is not None
    amax = hotpath_for(scheme, grid, bc).max_abs(scheme.velocity, 1)[0]
    return grid.dx_min / amax  # continuity/schemes.py:48-55


@dataclass(frozen=True, eq=False)
class FiniteVolumeScheme(ContinuityScheme, ConservationLawScheme):
    pass


@dataclass(frozen=True, eq=False)
class Godunov(FiniteVolumeScheme):
    """Upwind scheme (continuity/schemes.py:80-110)."""


@kernel_spec.register(Godunov)
def _spec(scheme: Godunov) -> dict:
    assert scheme.velocity is not None
    return {"equation": "continuity", "flux": "godunov", "alpha": 1.0, "velocity": scheme.velocity}


@numerical_flux.register(Godunov)
def _numerical_flux_continuity_godunov(scheme: Godunov, grid: Any, bc: Boundary, t: ScalarLike, u: Array) -> Array:
    assert scheme.velocity is not None
    assert scheme.rec is not None
    assert u.shape[-1] == grid.x.shape[0]
    from ..binding import NoBoundary

    return hotpath_for(scheme, grid, NoBoundary()).numerical_flux(u)
