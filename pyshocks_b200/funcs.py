"""Initial conditions / exact solutions used by the hot-path configurations
(``pyshocks/funcs.py:58-137, :311-357, :426-475``), on torch tensors."""

from __future__ import annotations

import math
from typing import Any

import torch

Array = torch.Tensor


def _f(mask: Array, like: Array) -> Array:
    return mask.to(like.dtype)


def ic_constant(grid: Any, x: Array, *, c: float = 1.0) -> Array:
    return torch.full_like(x, c)


def ic_sine(grid: Any, x: Array, *, k: float = 2) -> Array:
    # funcs.py:102-119
    xm = (grid.b + grid.a) / 2
    xh = (x - xm) / (grid.b - grid.a)
    return torch.sin(math.pi * k * xh)


def ic_sine_sine(grid: Any, x: Array, *, k1: float = 2, k2: float = 2) -> Array:
    # funcs.py:122-137
    xm = (grid.b + grid.a) / 2
    xh = (x - xm) / (grid.b - grid.a)
    return torch.sin(k1 * math.pi * xh + torch.sin(k2 * math.pi * xh) / math.pi)


def burgers_riemann(grid: Any, t: Any, x: Array, *, ul: float = 1.0, ur: float = 0.0, x0: float | None = None) -> Array:
    # funcs.py:311-357
    if x0 is None:
        x0 = 0.5 * (grid.a + grid.b)
    if not grid.a < x0 < grid.b:
        raise ValueError("'x0' must be in the domain [a, b].")
    if ul <= ur:
        h_l = _f(x < x0 + ul * t, x)
        h_c = _f(torch.logical_and(x0 + ul * t < x, x0 + ur * t > x), x)
        h_r = _f(x0 + ur * t < x, x)
        return ul * h_l + (x - x0) / (t + 1.0e-15) * h_c + ur * h_r
    s = (ul + ur) / 2.0
    h = _f(x < (x0 + s * t), x)
    return h * ul + (1 - h) * ur


def burgers_tophat(
    grid: Any, t: Any, x: Array, *, us: float = 0.0, uc: float = 1.0, xa: float | None = None, xb: float | None = None
) -> Array:
    # funcs.py:426-475
    xm = (grid.b + grid.a) / 2
    dx = grid.b - grid.a
    if xa is None:
        xa = xm - 0.25 * dx
    if xb is None:
        xb = xm + 0.25 * dx
    if xa >= xb:
        raise ValueError("Invalid sides (must be xa < xb).")
    if not grid.a < xa < grid.b:
        raise ValueError("'xa' must be in the domain [a, b].")
    if not grid.a < xb < grid.b:
        raise ValueError("'xb' must be in the domain [a, b].")
    if uc <= us:
        raise NotImplementedError("Inverse case with uc < us.")
    s = (uc + us) / 2
    h_l = _f(x < xa + us * t, x)
    h_e = _f(torch.logical_and(xa + us * t < x, x < xa + uc * t), x)
    h_c = _f(torch.logical_and(xa + uc * t < x, x < xb + s * t), x)
    h_r = _f(x > xb + s * t, x)
    return us * h_l + (x - xa) / (t + 1.0e-15) * h_e + uc * h_c + us * h_r
