"""Finite-volume schemes for the inviscid Burgers equation
(``pyshocks/burgers/schemes.py:30-256``): Godunov, Rusanov (LLF), LaxFriedrichs (global),
EngquistOsher, ESWENO32.  SSMUSCL / FluxSplitRusanov / SSWENO242 are different algorithms
and outside the hot path."""

from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Any

import torch

from .. import reconstruction
from ..binding import hotpath_for, kernel_spec
from ..schemes import (
    Boundary,
    ConservationLawScheme,
    SchemeBase,
    bind,
    flux,
    numerical_flux,
    predict_timestep,
)

Array = torch.Tensor
ScalarLike = Any


@dataclass(frozen=True, eq=False)
class BurgersScheme(SchemeBase):
    pass


@flux.register(BurgersScheme)
def _flux_burgers(scheme: BurgersScheme, t: ScalarLike, x: Array, u: Array) -> Array:
    return u**2 / 2  # burgers/schemes.py:37-39


def _interior_speed(scheme: SchemeBase, grid: Any, bc: Boundary, u: Array) -> Array:
    # jnp.max(jnp.abs(u[grid.i_])): warp-shuffle + block reduction kernel (psk_max_abs)
    smax = hotpath_for(scheme, grid, bc).max_abs(u, 1)
    return smax[0] if u.dim() == 1 else smax


@predict_timestep.register(BurgersScheme)
def _predict_timestep_burgers(scheme: BurgersScheme, grid: Any, bc: Boundary, t: ScalarLike, u: Array) -> Array:
    return 0.5 * grid.dx_min / _interior_speed(scheme, grid, bc, u)  # burgers/schemes.py:42-49


@dataclass(frozen=True, eq=False)
class FiniteVolumeScheme(BurgersScheme, ConservationLawScheme):
    pass


@dataclass(frozen=True, eq=False)
class Godunov(FiniteVolumeScheme):
    """Upwind flux with the averaged-speed switch (scalar_flux_upwind)."""


@dataclass(frozen=True, eq=False)
class Rusanov(FiniteVolumeScheme):
    """Rusanov / local Lax-Friedrichs flux."""

    alpha: float = 1.0


@dataclass(frozen=True, eq=False)
class LaxFriedrichs(Rusanov):
    """Global Lax-Friedrichs flux."""


@dataclass(frozen=True, eq=False)
class EngquistOsher(FiniteVolumeScheme):
    omega: float = field(default=0, init=False, repr=False)


@predict_timestep.register(Rusanov)
def _predict_timestep_burgers_rusanov(scheme: Rusanov, grid: Any, bc: Boundary, t: ScalarLike, u: Array) -> Array:
    # burgers/schemes.py:121-127
    return 0.5 * grid.dx_min ** (2 - scheme.alpha) / _interior_speed(scheme, grid, bc, u)


@dataclass(frozen=True, eq=False)
class ESWENO32(FiniteVolumeScheme):
    """Third-order Energy Stable WENO scheme (burgers/schemes.py:205-216): the upwind flux of
    the ESWENO32 reconstruction plus the dissipative flux built from its weights."""

    def __post_init__(self) -> None:
        if not isinstance(self.rec, reconstruction.ESWENO32):
            raise TypeError("ESWENO32 scheme requires the ESWENO32 reconstruction.")


@bind.register(ESWENO32)
def _bind_burgers_esweno32(scheme: ESWENO32, grid: Any, bc: Boundary) -> ESWENO32:
    # burgers/schemes.py:219-227: the parameters recommended by Carpenter (u0 = 1)
    from ..weno import es_weno_parameters

    eps, delta = es_weno_parameters(grid, torch.ones_like(grid.x))
    return replace(scheme, rec=replace(scheme.rec, eps=float(eps), delta=float(delta)))


@kernel_spec.register(ESWENO32)
def _spec_esweno32(scheme: ESWENO32) -> dict:
    return {"equation": "burgers", "flux": "esweno32", "alpha": 1.0, "velocity": None}


@kernel_spec.register(Godunov)
def _spec_godunov(scheme: Godunov) -> dict:
    return {"equation": "burgers", "flux": "godunov", "alpha": 1.0, "velocity": None}


@kernel_spec.register(Rusanov)
def _spec_rusanov(scheme: Rusanov) -> dict:
    return {"equation": "burgers", "flux": "rusanov", "alpha": float(scheme.alpha), "velocity": None}


@kernel_spec.register(LaxFriedrichs)
def _spec_lf(scheme: LaxFriedrichs) -> dict:
    return {"equation": "burgers", "flux": "lf", "alpha": float(scheme.alpha), "velocity": None}


@kernel_spec.register(EngquistOsher)
def _spec_eo(scheme: EngquistOsher) -> dict:
    return {"equation": "burgers", "flux": "eo", "alpha": 1.0, "velocity": None}


@numerical_flux.register(FiniteVolumeScheme)
def _numerical_flux_burgers(scheme: FiniteVolumeScheme, grid: Any, bc: Boundary, t: ScalarLike, u: Array) -> Array:
    # burgers/schemes.py:83-89, :110-118, :145-153, :188-196: u already carries its ghost cells
    from ..binding import NoBoundary

    return hotpath_for(scheme, grid, NoBoundary()).numerical_flux(u)
