"""Grids, quadrature and interior norms -- the host-side geometry the hot path needs.

Mirrors ``pyshocks/grid.py`` (Grid / UniformGrid fields :47-72, index helpers :83-117,
``make_uniform_cell_grid`` :131-179, Gauss-Legendre cell averages :343-435, interior-only
``norm`` / ``rnorm`` :444-531) with ``Array = torch.Tensor`` (CUDA, float64).  The grid arrays
are built with NumPy on the host exactly as the reference builds them and uploaded once.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable

import numpy as np
import torch

Array = torch.Tensor
SpatialFunction = Callable[[Array], Array]


def _device(device: Any = None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("pyshocks_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


@dataclass(frozen=True, eq=False)
class Grid:
    a: float
    """Left domain bound for [a, b]."""
    b: float
    """Right domain bound for [a, b]."""
    nghosts: int
    """Number of ghost cells."""
    x: Array
    """Solution point coordinates, shape ``(n + 2 g,)``."""
    dx: Array
    """Cell sizes."""
    f: Array
    """Midpoints between the solution points (faces), shape ``(n + 2 g + 1,)``."""
    df: Array
    """Cell sizes between the staggered points."""
    dx_min: Array
    dx_max: Array
    is_periodic: bool

    @property
    def dtype(self) -> torch.dtype:
        return self.x.dtype

    @property
    def device(self) -> torch.device:
        return self.x.device

    @property
    def n(self) -> int:
        # NOTE: like the reference (grid.py:78-81) this is x.size, ghosts INCLUDED
        return int(self.x.shape[0])

    @property
    def ncells(self) -> int:
        """Number of interior cells (not in the reference; its ``n`` includes ghosts)."""
        return int(self.x.shape[0]) - 2 * self.nghosts

    @property
    def i_(self) -> slice:
        return slice(self.nghosts, self.x.shape[0] - self.nghosts)

    @property
    def f_(self) -> slice:
        return slice(self.nghosts, self.f.shape[0] - self.nghosts)

    @property
    def b_(self) -> tuple[int, int, int]:
        return (-1, self.x.shape[0] - self.nghosts - 1, self.nghosts)

    @property
    def g_(self) -> tuple[None, slice, slice]:
        return (None, slice(self.x.shape[0] - self.nghosts, None), slice(None, self.nghosts))

    @property
    def gi_(self) -> tuple[None, slice, slice]:
        g = self.nghosts
        return (None, slice(self.x.shape[0] - 2 * g, self.x.shape[0] - g), slice(g, 2 * g))

    # host copies used to evaluate boundary data and Neumann offsets without a device sync
    @property
    def x_host(self) -> np.ndarray:
        cache = self.__dict__.get("_x_host")
        if cache is None:
            cache = self.x.detach().cpu().numpy()
            object.__setattr__(self, "_x_host", cache)
        return cache

    @property
    def h(self) -> float:
        cache = self.__dict__.get("_h")
        if cache is None:
            cache = float(self.dx_min)
            object.__setattr__(self, "_h", cache)
        return cache


@dataclass(frozen=True, eq=False)
class UniformGrid(Grid):
    pass


def make_uniform_cell_grid(
    a: float, b: float, n: int, *, nghosts: int = 1, dtype: Any = None, device: Any = None
) -> UniformGrid:
    """grid.py:131-179."""
    if b < a:
        raise ValueError(f"Incorrect interval a > b: {a!r} > {b!r}.")
    if n <= 0:
        raise ValueError(f"Number of cells should be > 0: {n!r} <= 0.")
    assert nghosts >= 0
    if dtype not in (None, torch.float64, np.float64):
        raise NotImplementedError("the hot path is fp64 only")
    dev = _device(device)
    h = (b - a) / n
    f = np.linspace(a - nghosts * h, b + nghosts * h, n + 2 * nghosts + 1, dtype=np.float64)
    x = (f[1:] + f[:-1]) / 2
    df = np.diff(x)
    dx = np.full_like(x, h)
    assert np.linalg.norm(np.diff(f) - h) < 1.0e-8 * h
    up = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(dev)  # noqa: E731
    grid = UniformGrid(
        a=a, b=b, nghosts=nghosts, x=up(x), dx=up(dx), f=up(f), df=up(df),
        dx_min=torch.tensor(h, dtype=torch.float64, device=dev),
        dx_max=torch.tensor(h, dtype=torch.float64, device=dev),
        is_periodic=False,
    )
    object.__setattr__(grid, "_x_host", x)
    object.__setattr__(grid, "_h", float(h))
    return grid


# {{{ cell averaging (grid.py:343-435)


@dataclass(frozen=True, eq=False)
class Quadrature:
    order: int
    x: Array
    """Quadrature points, shape ``(nnodes, ncells)``."""
    w: Array
    dx: Array

    def __post_init__(self) -> None:
        if self.x.shape != self.w.shape:
            raise ValueError(
                f"'x' and 'w' should have the same shape: got {self.x.shape} and {self.w.shape}."
            )

    @property
    def nnodes(self) -> int:
        return int(self.x.shape[0])

    @property
    def ncells(self) -> int:
        return int(self.x.shape[1])

    def __call__(self, fn: SpatialFunction, axis: int | None = None) -> Array:
        if axis not in {0, None}:
            raise ValueError(f"Unsupported axis value: {axis!r}.")
        val = fn(self.x) * self.w
        return torch.sum(val) if axis is None else torch.sum(val, dim=0)


def make_leggauss_quadrature(grid: Grid, order: int) -> Quadrature:
    if order < 1:
        raise ValueError(f"Invalid order: {order!r}.")
    return make_leggauss_quadrature_from_points(grid.f, order)


def make_leggauss_quadrature_from_points(x: Array, order: int) -> Quadrature:
    from numpy.polynomial.legendre import leggauss

    xi, wi = leggauss(order)
    xh = x.detach().cpu().numpy()
    dx = xh[1:] - xh[:-1]
    dxm = 0.5 * dx.reshape(1, -1)
    xm = 0.5 * (xh[1:] + xh[:-1]).reshape(1, -1)
    up = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(x.device)  # noqa: E731
    return Quadrature(order=order, x=up(xm + dxm * xi.reshape(-1, 1)), w=up(dxm * wi.reshape(-1, 1)), dx=up(dx))


def cell_average(quad: Quadrature, fn: SpatialFunction) -> Array:
    return quad(fn, axis=0) / quad.dx


# }}}

# {{{ norms (grid.py:444-531): interior only


def _norm(u: Array, dx: Array | float, p: Any) -> Array:
    u = torch.abs(u)
    if p == 1:
        return torch.sum(u * dx, dim=-1)
    if p == 2:
        return torch.sqrt(torch.sum(u**2 * dx, dim=-1))
    if p in {float("inf"), "inf"}:
        return torch.amax(u, dim=-1)
    if p in {-float("inf"), "-inf"}:
        return torch.amin(u, dim=-1)
    if p == "tvd":
        return torch.sum(torch.abs(torch.diff(u, dim=-1)), dim=-1)
    if isinstance(p, (int, float)):
        p = float(p)
        if p <= 0:
            raise ValueError(f"'p' must be a positive float: {p!r} <= 0.")
        return torch.sum(u**p * dx, dim=-1) ** (-1.0 / p)
    raise ValueError(f"Unrecognized norm order 'p': {p!r}.")


def norm(grid: Grid, u: float | Array, *, p: Any = 1, weighted: bool = False) -> Array:
    """Interior norm of *u*; a leading ensemble axis gives one value per row."""
    if isinstance(u, (int, float)) or u.dim() == 0:
        return torch.abs(torch.as_tensor(u, dtype=grid.x.dtype, device=grid.x.device))
    if u.shape[-1] == grid.x.shape[0]:
        dx = grid.dx[grid.i_] if weighted else 1.0
        return _norm(u[..., grid.i_], dx, p)
    if u.shape[-1] == grid.f.shape[0]:
        df = grid.df[grid.f_] if weighted else 1.0
        return _norm(u[..., grid.f_], df, p)
    raise ValueError(f"Array has unexpected shape: {tuple(u.shape)}")


def rnorm(
    grid: Grid, u: float | Array, v: float | Array, *, p: Any = 1, weighted: bool = False, atol: float = 1.0e-14
) -> Array:
    vnorm = norm(grid, v, p=p, weighted=weighted)
    vnorm = torch.where(vnorm < atol, torch.ones_like(vnorm), vnorm)
    return norm(grid, u - v, p=p, weighted=weighted) / vnorm


# }}}
