"""Ghost-cell boundary conditions and scalar numerical fluxes (``pyshocks/scalar.py``).

Boundary classes: :91-540 of the reference (``OneSidedBoundary``, ``TwoSidedBoundary``,
``DirichletBoundary``, ``NeumannBoundary``, ``PeriodicBoundary`` and the ``make_*`` helpers).
The SAT boundaries (:556-740) belong to the SBP family and are outside the hot path.
The flux functions keep the reference signatures and launch ``psk_numerical_flux``.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable

import torch

from .schemes import (
    Boundary,
    BoundaryType,
    ConservationLawScheme,
    apply_boundary,
    evaluate_boundary,
)

Array = torch.Tensor
ScalarLike = Any

# {{{ boundary conditions


@dataclass(frozen=True, eq=False)
class OneSidedBoundary(Boundary):
    side: int
    """``+1`` or ``-1``: the side on which this boundary condition applies."""


@dataclass(frozen=True, eq=False)
class TwoSidedBoundary(Boundary):
    left: OneSidedBoundary
    right: OneSidedBoundary

    def __post_init__(self) -> None:
        assert isinstance(self.left, OneSidedBoundary)
        assert isinstance(self.right, OneSidedBoundary)
        if self.left.side != -1:
            raise ValueError("Left boundary has incorrect side.")
        if self.right.side != +1:
            raise ValueError("Right boundary has incorrect side.")

    @property
    def boundary_type(self) -> BoundaryType:
        if self.left.boundary_type != self.right.boundary_type:
            raise NotImplementedError("Different boundaries on each side.")
        return self.left.boundary_type


@dataclass(frozen=True, eq=False)
class DirichletBoundary(OneSidedBoundary):
    g: Callable[[ScalarLike, Array], Array]
    """``g(t, x)`` evaluated at the ghost-cell centres of :attr:`side`."""

    @property
    def boundary_type(self) -> BoundaryType:
        return BoundaryType.Dirichlet


def make_dirichlet_boundary(ga: Callable, gb: Callable | None = None) -> TwoSidedBoundary:
    if gb is None:
        gb = ga
    return TwoSidedBoundary(left=DirichletBoundary(side=-1, g=ga), right=DirichletBoundary(side=+1, g=gb))


@dataclass(frozen=True, eq=False)
class NeumannBoundary(OneSidedBoundary):
    g: Callable[[ScalarLike], ScalarLike]
    """``g(t)``: the imposed normal derivative."""

    @property
    def boundary_type(self) -> BoundaryType:
        return BoundaryType.Neumann


def make_neumann_boundary(ga: Callable, gb: Callable | None = None) -> TwoSidedBoundary:
    if gb is None:
        gb = ga
    return TwoSidedBoundary(left=NeumannBoundary(side=-1, g=ga), right=NeumannBoundary(side=+1, g=gb))


@dataclass(frozen=True, eq=False)
class PeriodicBoundary(Boundary):
    @property
    def boundary_type(self) -> BoundaryType:
        return BoundaryType.Periodic


def _apply_boundary_kernel(bc: Boundary, grid: Any, t: ScalarLike, u: Array) -> Array:
    from .binding import boundary_path, ghost_data

    assert u.shape[-1] == grid.x.shape[0]
    if grid.nghosts == 0:
        return u.clone()
    hp = boundary_path(grid, bc)
    gd = ghost_data(bc, grid, t)
    if gd is not None:
        hp.set_ghost(gd)
    return hp.apply_boundary(u)


@apply_boundary.register(TwoSidedBoundary)
def _apply_boundary_two_sided(bc: TwoSidedBoundary, grid: Any, t: ScalarLike, u: Array) -> Array:
    # scalar.py:375-382 (left then right; here one launch fills both sides)
    if type(bc.left) is not type(bc.right):
        # different kinds on the two sides: the reference applies them one after the other here, and raises
        # NotImplementedError("Different boundaries on each side.") only where a scheme asks for
        # bc.boundary_type (scalar.py:369-370) -- i.e. in numerical_flux / apply_operator, as this package does
        return apply_boundary(bc.right, grid, t, apply_boundary(bc.left, grid, t, u))
    return _apply_boundary_kernel(bc, grid, t, u)


@apply_boundary.register(PeriodicBoundary)
def _apply_boundary_scalar_periodic(bc: PeriodicBoundary, grid: Any, t: ScalarLike, u: Array) -> Array:
    # scalar.py:529-540
    return _apply_boundary_kernel(bc, grid, t, u)


@apply_boundary.register(DirichletBoundary)
def _apply_boundary_scalar_dirichlet(bc: DirichletBoundary, grid: Any, t: ScalarLike, u: Array) -> Array:
    # scalar.py:418-427; one side only: plain tensor update (host-level convenience path)
    assert u.shape[-1] == grid.x.shape[0]
    ito = grid.g_[bc.side]
    out = u.clone()
    out[..., ito] = torch.as_tensor(bc.g(t, grid.x[ito]), dtype=u.dtype, device=u.device)
    return out


@apply_boundary.register(NeumannBoundary)
def _apply_boundary_scalar_neumann(bc: NeumannBoundary, grid: Any, t: ScalarLike, u: Array) -> Array:
    # scalar.py:472-500; one side only
    assert u.shape[-1] == grid.x.shape[0]
    g = grid.nghosts
    ifrom = torch.arange(*grid.gi_[bc.side].indices(u.shape[-1]), device=u.device)
    if bc.side == -1:
        ito = torch.arange(g - 1, -1, -1, device=u.device)
    else:
        ito = torch.arange(u.shape[-1] - 1, u.shape[-1] - g - 1, -1, device=u.device)
    out = u.clone()
    out[..., ito] = u[..., ifrom] + bc.side * (grid.x[ifrom] - grid.x[ito]) * float(bc.g(t))
    return out


@evaluate_boundary.register(TwoSidedBoundary)
def _evaluate_boundary_two_sided(bc: TwoSidedBoundary, grid: Any, t: ScalarLike, u: Array) -> Array:
    return evaluate_boundary(bc.left, grid, t, u) + evaluate_boundary(bc.right, grid, t, u)


@evaluate_boundary.register(PeriodicBoundary)
def _evaluate_boundary_scalar_periodic(bc: PeriodicBoundary, grid: Any, t: ScalarLike, u: Array) -> Array:
    return torch.zeros_like(u)


# }}}

# {{{ fluxes (scalar.py:91-322): reference signatures, kernel implementations


def _flux_via_kernel(scheme: ConservationLawScheme, grid: Any, u: Array, flux_name: str, alpha: float = 1.0) -> Array:
    from .binding import NoBoundary, hotpath_for, kernel_spec

    spec = kernel_spec(scheme)
    if spec["flux"] != flux_name or abs(spec["alpha"] - alpha) > 0.0:
        raise NotImplementedError(
            f"{type(scheme).__name__} is bound to the '{spec['flux']}' flux (alpha = {spec['alpha']}); "
            f"call numerical_flux(scheme, ...) or use a scheme of the '{flux_name}' family"
        )
    assert u.shape[-1] == grid.x.shape[0]
    assert scheme.rec is not None
    return hotpath_for(scheme, grid, NoBoundary()).numerical_flux(u)


def scalar_flux_upwind(scheme: ConservationLawScheme, grid: Any, bc: BoundaryType, t: ScalarLike, a: Array, u: Array) -> Array:
    """Averaged-speed upwind switch (scalar.py:91-132); ``a`` must be the scheme's own speed."""
    return _flux_via_kernel(scheme, grid, u, "godunov")


def scalar_flux_rusanov(
    scheme: ConservationLawScheme, grid: Any, bc: BoundaryType, t: ScalarLike, a: Array, u: Array, alpha: ScalarLike = 1.0
) -> Array:
    """Rusanov / local Lax-Friedrichs flux (scalar.py:192-249)."""
    return _flux_via_kernel(scheme, grid, u, "rusanov", float(alpha))


def scalar_flux_lax_friedrichs(
    scheme: ConservationLawScheme, grid: Any, bc: BoundaryType, t: ScalarLike, a: Array, u: Array, alpha: ScalarLike = 1.0
) -> Array:
    """Global Lax-Friedrichs flux (scalar.py:258-278)."""
    return _flux_via_kernel(scheme, grid, u, "lf", float(alpha))


def scalar_flux_engquist_osher(
    scheme: ConservationLawScheme, grid: Any, bc: BoundaryType, t: ScalarLike, a: Array, u: Array, omega: ScalarLike = 0.0
) -> Array:
    """Engquist-Osher flux for convex fluxes, ``omega = 0`` (scalar.py:287-322)."""
    if float(omega) != 0.0:
        raise NotImplementedError("only omega = 0 (Burgers) is on the hot path")
    return _flux_via_kernel(scheme, grid, u, "eo")


# }}}
