"""The operator API of pyshocks (``pyshocks/schemes.py``): scheme base classes, boundary
base classes and the ``functools.singledispatch`` generic functions every equation package
registers against.  This is the drop-in boundary of the hot path:
names, argument meaning and error behaviour follow the reference; the registered
implementations launch the kernels of ``libpsk.so`` instead of tracing JAX.
"""

from __future__ import annotations

import enum
from abc import ABC, abstractmethod
from dataclasses import dataclass
from functools import singledispatch
from typing import TYPE_CHECKING, Any, TypeVar

import torch

if TYPE_CHECKING:
    from .grid import Grid
    from .reconstruction import Reconstruction

Array = torch.Tensor
ScalarLike = Any

# {{{ schemes (schemes.py:78-190)


@dataclass(frozen=True, eq=False)
class SchemeBase:
    rec: "Reconstruction"
    """Reconstruction used to get high-order face values."""

    @property
    def name(self) -> str:
        return f"{type(self).__name__}_{self.rec.name}".lower()

    @property
    def order(self) -> int:
        return self.rec.order

    @property
    def stencil_width(self) -> int:
        return self.rec.stencil_width


SchemeT = TypeVar("SchemeT", bound=SchemeBase)


@singledispatch
def bind(scheme: SchemeT, grid: "Grid", bc: "Boundary") -> SchemeT:
    """schemes.py:124-139."""
    if isinstance(scheme, SchemeBase):
        return scheme
    raise NotImplementedError(type(scheme).__name__)


@singledispatch
def apply_operator(scheme: SchemeBase, grid: "Grid", bc: "Boundary", t: ScalarLike, u: Array) -> Array:
    """Method-of-lines right-hand side at every cell of ``grid.x`` (schemes.py:142-166)."""
    raise NotImplementedError(type(scheme).__name__)


@singledispatch
def predict_timestep(scheme: SchemeBase, grid: "Grid", bc: "Boundary", t: ScalarLike, u: Array) -> Array:
    """schemes.py:169-190."""
    raise NotImplementedError(type(scheme).__name__)


@dataclass(frozen=True, eq=False)
class FiniteVolumeSchemeBase(SchemeBase):
    pass


@dataclass(frozen=True, eq=False)
class FiniteDifferenceSchemeBase(SchemeBase):
    pass


# }}}

# {{{ conservation laws (schemes.py:287-346)


@dataclass(frozen=True, eq=False)
class ConservationLawScheme(FiniteVolumeSchemeBase):
    pass


@singledispatch
def flux(scheme: SchemeBase, t: ScalarLike, x: Array, u: Array) -> Array:
    """Physical flux (schemes.py:305-318)."""
    raise NotImplementedError(type(scheme).__name__)


@singledispatch
def numerical_flux(scheme: SchemeBase, grid: "Grid", bc: "Boundary", t: ScalarLike, u: Array) -> Array:
    """Flux at every face of ``grid.f`` (schemes.py:321-336)."""
    raise NotImplementedError(type(scheme).__name__)


@apply_operator.register(ConservationLawScheme)
def _apply_operator_conservation_law(
    scheme: ConservationLawScheme, grid: "Grid", bc: "Boundary", t: ScalarLike, u: Array
) -> Array:
    # schemes.py:339-346: apply_boundary -> numerical_flux -> -(f[1:] - f[:-1]) / dx, as ONE
    # fused launch (plus one tiny launch for the ghost rows the reference also produces)
    from .binding import hotpath_for

    from .timestepping import _trace_apply_operator

    hp = hotpath_for(scheme, grid, bc, t)
    out = hp.apply_operator(u)
    _trace_apply_operator(scheme, grid, bc, t, u, out)
    return out


# }}}

# {{{ boundary conditions (schemes.py:393-456)


@enum.unique
class BoundaryType(enum.Enum):
    Periodic = enum.auto()
    Dirichlet = enum.auto()
    Neumann = enum.auto()
    HomogeneousNeumann = enum.auto()


@dataclass(frozen=True, eq=False)
class Boundary(ABC):
    @property
    @abstractmethod
    def boundary_type(self) -> BoundaryType:
        """Broad class of the boundary condition."""


@singledispatch
def apply_boundary(bc: Boundary, grid: "Grid", t: ScalarLike, u: Array) -> Array:
    """Copy of *u* with its ghost layer set (schemes.py:431-442)."""
    raise NotImplementedError(type(bc).__name__)


@singledispatch
def evaluate_boundary(bc: Boundary, grid: "Grid", t: ScalarLike, u: Array) -> Array:
    """schemes.py:445-456."""
    raise NotImplementedError(type(bc).__name__)


# }}}
