"""WENO-JS stencil tables (``pyshocks/weno.py:54-98, :166-244``).

In the reference these tables drive 18 small convolutions per right-hand side; here they are
compiled into the kernels (``csrc/psk_math.cuh``).  The tables are kept for API parity and for
the test that pins the kernels' coefficients against them
(``tests/test_gpu_kernels.py::test_kernel_coefficients_are_those_of_the_stencil_tables``).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Stencil:
    a: np.ndarray | None
    """Weights of the squared terms of the smoothness indicators."""
    b: np.ndarray | None
    """Stencils of the smoothness indicators."""
    c: np.ndarray
    """Interpolation stencils."""
    d: np.ndarray
    """Ideal weights."""


def weno_js_32_coefficients() -> Stencil:
    # weno.py:166-203 (taps listed i+1, i, i-1)
    a = np.array([1.0])
    b = np.array([[[0.0, 1.0, -1.0]], [[1.0, -1.0, 0.0]]])
    c = np.array([[0.0, 3.0 / 2.0, -1.0 / 2.0], [1.0 / 2.0, 1.0 / 2.0, 0.0]])
    d = np.array([[1.0 / 3.0, 2.0 / 3.0]]).T
    return Stencil(a=a, b=b, c=c, d=d)


def weno_js_53_coefficients() -> Stencil:
    # weno.py:206-244 (taps listed i+2, i+1, i, i-1, i-2)
    a = np.array([13.0 / 12.0, 1.0 / 4.0])
    b = np.array(
        [
            [[0.0, 0.0, 1.0, -2.0, 1.0], [0.0, 0.0, 3.0, -4.0, 1.0]],
            [[0.0, 1.0, -2.0, 1.0, 0.0], [0.0, 1.0, 0.0, -1.0, 0.0]],
            [[1.0, -2.0, 1.0, 0.0, 0.0], [1.0, -4.0, 3.0, 0.0, 0.0]],
        ]
    )
    c = np.array(
        [
            [0.0, 0.0, 11.0 / 6.0, -7.0 / 6.0, 2.0 / 6.0],
            [0.0, 2.0 / 6.0, 5.0 / 6.0, -1.0 / 6.0, 0.0],
            [-1.0 / 6.0, 5.0 / 6.0, 2.0 / 6.0, 0.0, 0.0],
        ]
    )
    d = np.array([[1.0 / 10.0, 6.0 / 10.0, 3.0 / 10.0]]).T
    return Stencil(a=a, b=b, c=c, d=d)


def es_weno_parameters(grid, u0):  # noqa: ANN001, ANN201
    """``(eps, delta)`` of the ESWENO32 scheme from the grid and the initial condition
    (weno.py:263-281, Equations 65-66 of Yamaleev & Carpenter 2009).  Evaluated on the host
    (a one-off scalar at bind time), in the reference's expression order."""
    import torch

    i = grid.i_
    u0h = u0.detach().cpu().numpy() if isinstance(u0, torch.Tensor) else np.asarray(u0, dtype=np.float64)
    dx = np.full_like(u0h, grid.h)
    dx_min = np.float64(grid.h)
    eps = np.sum(dx[i] * np.abs(u0h[i])) * dx_min**2
    delta = dx_min**2
    return eps, delta
