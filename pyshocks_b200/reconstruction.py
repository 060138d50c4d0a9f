"""Reconstruction classes and the ``reconstruct`` generic function
(``pyshocks/reconstruction.py:79-163, :305-377, :528-564``).

In scope: ``ConstantReconstruction``, ``WENOJS32``, ``WENOJS53``, ``ESWENO32``
(``reconstruction.py:386-439``).  MUSCL / MUSCLS / SSWENO242 are other
scheme families and raise ``ValueError`` from the name registry like an
unknown name does in the reference.
"""

from __future__ import annotations

from dataclasses import dataclass, fields
from functools import singledispatch
from typing import Any, ClassVar

import torch

from . import weno
from .schemes import BoundaryType

Array = torch.Tensor


@dataclass(frozen=True)
class Reconstruction:
    @property
    def name(self) -> str:
        return type(self).__name__.lower()

    @property
    def order(self) -> int:
        raise NotImplementedError(type(self).__name__)

    @property
    def stencil_width(self) -> int:
        raise NotImplementedError(type(self).__name__)


@singledispatch
def reconstruct(rec: Reconstruction, grid: Any, bc: BoundaryType, f: Array, u: Array, wavespeed: Array) -> tuple[Array, Array]:
    """``(fl, fr)``: values at the left / right face of every cell (reconstruction.py:79-118)."""
    raise NotImplementedError(type(rec).__name__)


@dataclass(frozen=True)
class ConstantReconstruction(Reconstruction):
    @property
    def name(self) -> str:
        return "constant"

    @property
    def order(self) -> int:
        return 1

    @property
    def stencil_width(self) -> int:
        return 1


@reconstruct.register(ConstantReconstruction)
def _reconstruct_first_order(rec: ConstantReconstruction, grid: Any, bc: BoundaryType, f: Array, u: Array, wavespeed: Array) -> tuple[Array, Array]:
    assert grid.nghosts >= rec.stencil_width
    return f, f


@dataclass(frozen=True)
class WENOJS(Reconstruction):
    eps: float
    s: ClassVar[weno.Stencil]

    @property
    def stencil_width(self) -> int:
        return self.order


@dataclass(frozen=True)
class WENOJS32(WENOJS):
    eps: float = 1.0e-6

    def __post_init__(self) -> None:
        object.__setattr__(self, "s", weno.weno_js_32_coefficients())

    @property
    def order(self) -> int:
        return 2


@dataclass(frozen=True)
class WENOJS53(WENOJS):
    eps: float = 1.0e-12

    def __post_init__(self) -> None:
        object.__setattr__(self, "s", weno.weno_js_53_coefficients())

    @property
    def order(self) -> int:
        return 3


def _reconstruct_launch(rec: Any, grid: Any, f: Array, what: str) -> tuple[Array, Array]:
    # one launch of psk_reconstruct (zero padded like jnp.convolve "same")
    from . import config
    from .grid import UniformGrid
    from .path import HotPath

    assert grid.nghosts >= rec.stencil_width
    if not isinstance(grid, UniformGrid):
        raise NotImplementedError(f"{what} is only implemented for uniform grids.")
    cache = grid.__dict__.setdefault("_psk_rec", {}) if hasattr(grid, "__dict__") else {}
    eps = float(rec.eps)
    key = (rec.name, eps, config.MATH)
    hp = cache.get(key)
    if hp is None:
        hp = HotPath(equation="burgers", flux="rusanov", rec=rec.name, bc="none",
                     n=grid.x.shape[0] - 2 * grid.nghosts, g=grid.nghosts, dx=grid.h, eps=eps,
                     math=config.MATH, device=grid.x.device)
        cache[key] = hp
    return hp.reconstruct(f)


@reconstruct.register(WENOJS)
def _reconstruct_wenojs(rec: WENOJS, grid: Any, bc: BoundaryType, f: Array, u: Array, wavespeed: Array) -> tuple[Array, Array]:
    # reconstruction.py:358-377
    return _reconstruct_launch(rec, grid, f, "WENO-JS")


@dataclass(frozen=True)
class ESWENO32(Reconstruction):
    """Third-order WENO reconstruction with the modified weights of the Energy Stable WENO
    scheme (reconstruction.py:386-410): ``alpha_k = d_k (1 + tau / (eps + beta_k))`` on the
    JS-3 stencils (weno.py:284-296)."""

    eps: Any = 1.0e-6
    delta: Any = 1.0e-6
    s: ClassVar[weno.Stencil]

    def __post_init__(self) -> None:
        object.__setattr__(self, "s", weno.weno_js_32_coefficients())

    @property
    def order(self) -> int:
        return 2

    @property
    def stencil_width(self) -> int:
        return 2


@reconstruct.register(ESWENO32)
def _reconstruct_esweno32(rec: ESWENO32, grid: Any, bc: BoundaryType, f: Array, u: Array, wavespeed: Array) -> tuple[Array, Array]:
    # reconstruction.py:420-439
    return _reconstruct_launch(rec, grid, f, "ES-WENO")


_RECONSTRUCTION: dict[str, type[Reconstruction]] = {
    "default": ConstantReconstruction,
    "constant": ConstantReconstruction,
    "wenojs32": WENOJS32,
    "wenojs53": WENOJS53,
    "esweno32": ESWENO32,
}


def reconstruction_ids() -> tuple[str, ...]:
    return tuple(_RECONSTRUCTION.keys())


def make_reconstruction_from_name(name: str, **kwargs: Any) -> Reconstruction:
    cls = _RECONSTRUCTION.get(name)
    if cls is None:
        raise ValueError(
            f"Reconstruction {name!r} not found (outside the WENO-JS hot path?). "
            f"Try one of {', '.join(reconstruction_ids())}."
        )
    return cls(**{f.name: kwargs[f.name] for f in fields(cls) if f.name in kwargs})
