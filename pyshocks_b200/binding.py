"""Glue between the pyshocks-shaped objects (scheme, grid, boundary) and :class:`HotPath`.

``hotpath_for(scheme, grid, bc, t)`` returns the cached kernel binding of a scheme on a grid
with the boundary data for time ``t`` loaded (Dirichlet values ``g(t, x_ghost)`` or Neumann
offsets), so that the registered ``apply_operator`` / ``numerical_flux`` / ``advance``
implementations are one or three kernel launches.
"""

from __future__ import annotations

from functools import singledispatch
from typing import Any

import numpy as np
import torch

from .path import HotPath
from .schemes import Boundary, SchemeBase


@singledispatch
def kernel_spec(scheme: SchemeBase) -> dict[str, Any]:
    """``{"equation", "flux", "alpha", "velocity"}`` of a scheme the kernels implement.

    Equation packages register their in-scope scheme classes; anything else (ESWENO, SBP-SAT,
    flux splitting, MUSCL, ...) is outside the hot path and raises like an unregistered type
    does in the reference (schemes.py:139, :166, :190)."""
    raise NotImplementedError(type(scheme).__name__)


def boundary_kind(bc: Boundary) -> str:
    from .scalar import DirichletBoundary, NeumannBoundary, PeriodicBoundary, TwoSidedBoundary

    if isinstance(bc, PeriodicBoundary):
        return "periodic"
    if isinstance(bc, TwoSidedBoundary):
        if isinstance(bc.left, DirichletBoundary) and isinstance(bc.right, DirichletBoundary):
            return "dirichlet"
        if isinstance(bc.left, NeumannBoundary) and isinstance(bc.right, NeumannBoundary):
            return "neumann"
        raise NotImplementedError("Different boundaries on each side.")  # scalar.py:369-370
    if isinstance(bc, NoBoundary):
        return "none"
    raise NotImplementedError(type(bc).__name__)


class NoBoundary(Boundary):
    """Ghost cells are taken as found (halo-exchanged slabs of a decomposed grid)."""

    @property
    def boundary_type(self):  # noqa: ANN201
        from .schemes import BoundaryType

        return BoundaryType.Dirichlet


def _host(v: Any, n: int) -> np.ndarray:
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    return np.broadcast_to(np.asarray(v, dtype=np.float64), (n,))


def ghost_data(bc: Boundary, grid: Any, t: Any) -> torch.Tensor | np.ndarray | None:
    """The ``2 g`` numbers per row the kernels need for time ``t`` (include/psk.h, psk_bc)."""
    from .scalar import DirichletBoundary, NeumannBoundary, TwoSidedBoundary

    if not isinstance(bc, TwoSidedBoundary):
        return None
    g, nx = grid.nghosts, grid.x.shape[0]
    if isinstance(bc.left, DirichletBoundary):
        # scalar.py:424-425: bc.g(t, grid.x[ghost]) per side
        left = bc.left.g(t, grid.x[:g])
        right = bc.right.g(t, grid.x[nx - g :])
        if isinstance(left, torch.Tensor) and isinstance(right, torch.Tensor) and left.is_cuda:
            return torch.cat([left.reshape(-1).expand(g) if left.numel() == 1 else left.reshape(-1),
                              right.reshape(-1).expand(g) if right.numel() == 1 else right.reshape(-1)])
        return np.concatenate([_host(left, g), _host(right, g)])
    if isinstance(bc.left, NeumannBoundary):
        # scalar.py:490-498: ub = u[ifrom] + side * (x[ifrom] - x[ito]) * g(t)
        x = grid.x_host
        ga, gb = float(bc.left.g(t)), float(bc.right.g(t))
        out = np.empty(2 * g)
        for k in range(g):
            out[k] = (-1) * (x[2 * g - 1 - k] - x[k]) * ga
            out[g + k] = (+1) * (x[nx - g - 1 - k] - x[nx - g + k]) * gb
        return out
    return None


def _paths(obj: Any) -> dict:
    cache = obj.__dict__.get("_psk_paths")
    if cache is None:
        cache = {}
        object.__setattr__(obj, "_psk_paths", cache)
    return cache


def hotpath_for(scheme: SchemeBase, grid: Any, bc: Boundary, t: Any = None, *, math: str | None = None) -> HotPath:
    from . import config

    math = config.MATH if math is None else math
    kind = boundary_kind(bc)
    key = (id(grid), kind, math)
    cache = _paths(scheme)
    entry = cache.get(key)
    # the entry keeps the grid alive (so its id cannot be reused) and is only trusted for that very grid
    hp = entry[1] if (entry is not None and entry[0] is grid) else None
    if hp is None:
        from .grid import UniformGrid

        spec = kernel_spec(scheme)
        if not isinstance(grid, UniformGrid):
            # reconstruction.py:371-372
            raise NotImplementedError("WENO-JS is only implemented for uniform grids.")
        assert grid.nghosts >= scheme.rec.stencil_width  # reconstruction.py:369, :161
        nu = None
        if abs(spec["alpha"] - 1.0) > 1.0e-8:
            nu = np.diff(grid.x_host) ** (spec["alpha"] - 1)  # grid.df ** (alpha - 1), scalar.py:231-232
        hp = HotPath(
            equation=spec["equation"], flux=spec["flux"], rec=scheme.rec.name, bc=kind,
            n=grid.x.shape[0] - 2 * grid.nghosts, g=grid.nghosts, dx=grid.h, eps=float(getattr(scheme.rec, "eps", 0.0)),
            math=math, nu=nu, velocity=spec["velocity"], device=grid.x.device,
            delta=float(getattr(scheme.rec, "delta", 0.0)),
        )
        cache[key] = (grid, hp)
    if t is not None:
        gd = ghost_data(bc, grid, t)
        if gd is not None:
            hp.set_ghost(gd)
    return hp


def boundary_path(grid: Any, bc: Boundary) -> HotPath:
    """Kernel binding used by ``apply_boundary`` alone (no scheme involved)."""
    kind = boundary_kind(bc)
    cache = _paths(grid)
    hp = cache.get(kind)
    if hp is None:
        hp = HotPath(equation="burgers", flux="rusanov", rec="constant", bc=kind,
                     n=grid.x.shape[0] - 2 * grid.nghosts, g=max(grid.nghosts, 1), dx=grid.h, eps=0.0,
                     device=grid.x.device)
        cache[kind] = hp
    return hp
