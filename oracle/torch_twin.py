"""Differentiable twin (torch fp64, CPU) of ``oracle/pyshocks_oracle.py``.

TEST INFRASTRUCTURE ONLY (same import rules as the NumPy oracle).  The
reference obtains its discrete adjoint by automatic differentiation of
``advance`` (``jax.jacfwd``, timestepping.py:174, then ``jac.T @ p`` at
:205-206).  JAX is unavailable here, so this module restates the same forward
functions with torch ops and lets ``torch.autograd`` stand in for JAX autodiff.
The derivative conventions at kinks agree with JAX: ``where`` differentiates the
selected branch only, ``maximum``/``minimum`` split ties 1/2-1/2, ``abs'(0)=0``,
a full ``max`` reduction shares the gradient between tied positions.

Pinning: forward values are checked against the NumPy oracle, and the
``adjoint_step`` results against the golden vectors recorded from the
reference's own ``adjoint_step`` (``tests/golden/adjoint_*.npz``).
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import torch

from . import pyshocks_oracle as po

_F64 = torch.float64


def _t(x: object) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=_F64)


def _conv_same(u: torch.Tensor, w: np.ndarray) -> torch.Tensor:
    """``numpy.convolve(u, w, "same")`` for an odd-length kernel, zero padded,
    accumulated in ascending memory order like NumPy does (convolve.py:113-114)."""
    r = len(w) // 2
    up = torch.nn.functional.pad(u, (r, r))
    n = u.shape[0]
    acc = None
    for k in range(len(w)):
        coef = float(w[len(w) - 1 - k])
        if coef == 0.0:
            continue
        term = up[k : k + n] * coef
        acc = term if acc is None else acc + term
    return acc


def _weno_js_side(s: dict, eps: float, f: torch.Tensor) -> torch.Tensor:
    # weno.py:134-140, 157, 253-256; reconstruction.py:351-355
    nst = s["b"].shape[0]
    beta = []
    for i in range(nst):
        acc = None
        for j in range(s["a"].size):
            term = float(s["a"][j]) * _conv_same(f, s["b"][i, j, :]) ** 2
            acc = term if acc is None else acc + term
        beta.append(acc)
    alpha = [float(s["d"][i, 0]) / (eps + beta[i]) ** 2 for i in range(nst)]
    total = alpha[0]
    for i in range(1, nst):
        total = total + alpha[i]
    out = None
    for i in range(nst):
        term = (alpha[i] / total) * _conv_same(f, s["c"][i, :])
        out = term if out is None else out + term
    return out


def _smoothness(s: dict, f: torch.Tensor) -> list[torch.Tensor]:
    # weno.py:134-140
    beta = []
    for i in range(s["b"].shape[0]):
        acc = None
        for j in range(s["a"].size):
            term = float(s["a"][j]) * _conv_same(f, s["b"][i, j, :]) ** 2
            acc = term if acc is None else acc + term
        beta.append(acc)
    return beta


def es_weno_weights(s: dict, f: torch.Tensor, eps: float) -> list[torch.Tensor]:
    # weno.py:284-296: alpha_k = d_k (1 + tau / (eps + beta_k)), tau zero at the two ends of the array
    beta = _smoothness(s, f)
    tau = torch.nn.functional.pad((f[2:] - 2 * f[1:-1] + f[:-2]) ** 2, (1, 1))
    alpha = [float(s["d"][i, 0]) * (1 + tau / (eps + beta[i])) for i in range(len(beta))]
    total = alpha[0]
    for i in range(1, len(alpha)):
        total = total + alpha[i]
    return [a / total for a in alpha]


def _es_weno_side(s: dict, eps: float, f: torch.Tensor) -> torch.Tensor:
    # reconstruction.py:413-417
    omega = es_weno_weights(s, f, eps)
    out = None
    for i in range(len(omega)):
        term = omega[i] * _conv_same(f, s["c"][i, :])
        out = term if out is None else out + term
    return out


def reconstruct(rec: po.Reconstruction, f: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    if rec.name == "constant":
        return f, f
    if rec.name == "esweno32":  # reconstruction.py:420-439
        fr = _es_weno_side(po._JS32, rec.eps, f)
        fl = torch.flip(_es_weno_side(po._JS32, rec.eps, torch.flip(f, (0,))), (0,))
        return fl, fr
    s = po._JS53 if rec.name == "wenojs53" else po._JS32
    fr = _weno_js_side(s, rec.eps, f)
    fl = torch.flip(_weno_js_side(s, rec.eps, torch.flip(f, (0,))), (0,))
    return fl, fr


def apply_boundary(bc: object, grid: po.OracleGrid, t: float, u: torch.Tensor) -> torch.Tensor:
    g, nx = grid.g, grid.nx
    if isinstance(bc, po.Periodic):
        return torch.cat([u[nx - 2 * g : nx - g], u[g : nx - g], u[g : 2 * g]])
    if isinstance(bc, po.Dirichlet):
        gb = bc.gb if bc.gb is not None else bc.ga
        left = _t(bc.ga(t, grid.x[:g]) + np.zeros(g))
        right = _t(gb(t, grid.x[nx - g :]) + np.zeros(g))
        return torch.cat([left, u[g : nx - g], right])
    if isinstance(bc, po.Neumann):
        gb = bc.gb if bc.gb is not None else bc.ga
        x = _t(grid.x)
        ifrom = torch.arange(g, 2 * g)
        ito = torch.arange(g - 1, -1, -1)
        left = u[ifrom] + (-1) * (x[ifrom] - x[ito]) * float(bc.ga(t))
        ifrom_r = torch.arange(nx - 2 * g, nx - g)
        ito_r = torch.arange(nx - 1, nx - g - 1, -1)
        right = u[ifrom_r] + (+1) * (x[ifrom_r] - x[ito_r]) * float(gb(t))
        # left[k] goes to index ito[k] = g-1-k -> reversed order in memory
        return torch.cat([torch.flip(left, (0,)), u[g : nx - g], torch.flip(right, (0,))])
    if isinstance(bc, po.NoBoundary):
        return u
    raise NotImplementedError(type(bc).__name__)


def _physical_flux(scheme: po.Scheme, u: torch.Tensor) -> torch.Tensor:
    if scheme.equation == "burgers":
        return u**2 / 2
    if scheme.equation == "continuity":
        return _t(scheme.velocity) * u
    raise NotImplementedError(scheme.equation)


def numerical_flux(scheme: po.Scheme, grid: po.OracleGrid, u: torch.Tensor) -> torch.Tensor:
    pad = torch.nn.functional.pad
    rec = scheme.rec
    if scheme.equation == "burgers":
        ul, ur = reconstruct(rec, u)
        if scheme.flux in ("godunov", "esweno32"):
            fl, fr = _physical_flux(scheme, ul), _physical_flux(scheme, ur)
            aavg = (ur[:-1] + ul[1:]) / 2
            fnum = pad(torch.where(aavg > 0, fr[:-1], fl[1:]), (1, 1))
            if scheme.flux == "esweno32":  # burgers/schemes.py:230-256: + the dissipative flux of ESWENO
                omega = es_weno_weights(po._JS32, u, rec.eps)[0]
                dom = omega[1:] - omega[:-1]
                mu = torch.sqrt(dom**2 + rec.delta**2) / 8.0
                fnum = fnum + pad(-(mu + dom / 8.0) * (u[1:] - u[:-1]), (1, 1))
            return fnum
        if scheme.flux in ("rusanov", "lf"):
            if abs(scheme.alpha - 1.0) > 1.0e-8:
                nu = _t(grid.df ** (scheme.alpha - 1))
            else:
                nu = torch.tensor(1.0, dtype=_F64)
            fl, fr = _physical_flux(scheme, ul), _physical_flux(scheme, ur)
            if scheme.flux == "lf":
                a = torch.amax(torch.abs(u))
            else:
                a = torch.abs(u)
                a = torch.maximum(a[1:], a[:-1])
            fnum = 0.5 * (fl[1:] + fr[:-1]) - 0.5 * a * nu * (ul[1:] - ur[:-1])
            return pad(fnum, (1, 1))
        if scheme.flux == "eo":
            zero = torch.zeros((), dtype=_F64)
            fr = _physical_flux(scheme, torch.maximum(ur, zero))
            fl = _physical_flux(scheme, torch.minimum(ul, zero))
            return pad(fr[:-1] + fl[1:] - 0.0, (1, 1))
        raise NotImplementedError(scheme.flux)
    a = _t(scheme.velocity)
    ul, ur = reconstruct(rec, u)
    al, ar = reconstruct(rec, a)
    aavg = (ar[:-1] + al[1:]) / 2
    if scheme.equation == "advection":
        return pad(torch.where(aavg > 0, ur[:-1], ul[1:]), (1, 1))
    if scheme.equation == "continuity":
        return pad(torch.where(aavg > 0, ar[:-1] * ur[:-1], al[1:] * ul[1:]), (1, 1))
    raise NotImplementedError(scheme.equation)


def apply_operator(scheme: po.Scheme, grid: po.OracleGrid, bc: object, t: float, u: torch.Tensor) -> torch.Tensor:
    w = apply_boundary(bc, grid, t, u)
    f = numerical_flux(scheme, grid, w)
    dx = _t(grid.dx)
    if scheme.equation == "advection":
        return -_t(scheme.velocity) * (f[1:] - f[:-1]) / dx
    return -(f[1:] - f[:-1]) / dx


def ssprk33_advance(source: Callable, dt: float, t: float, u: torch.Tensor) -> torch.Tensor:
    k1 = u + dt * source(t, u)
    k2 = 3.0 / 4.0 * u + 1.0 / 4.0 * (k1 + dt * source(t + dt, k1))
    return 1.0 / 3.0 * u + 2.0 / 3.0 * (k2 + dt * source(t + 0.5 * dt, k2))


def rhs_vjp(scheme: po.Scheme, grid: po.OracleGrid, bc: object, t: float, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """``J_L(u)^T v`` for the RHS ``L = apply_operator`` (all ``nx`` rows)."""
    ut = _t(u).clone().requires_grad_(True)
    out = apply_operator(scheme, grid, bc, t, ut)
    (g,) = torch.autograd.grad(out, ut, grad_outputs=_t(v), allow_unused=True)
    return np.zeros_like(u) if g is None else g.numpy()


def step_vjp(scheme: po.Scheme, grid: po.OracleGrid, bc: object, dt: float, t: float, u: np.ndarray, p: np.ndarray) -> np.ndarray:
    """``(d advance / d u)^T p`` for one full SSPRK33 step (timestepping.py:205-206)."""
    ut = _t(u).clone().requires_grad_(True)

    def source(tt: float, x: torch.Tensor) -> torch.Tensor:
        return apply_operator(scheme, grid, bc, tt, x)

    out = ssprk33_advance(source, float(dt), float(t), ut)
    (g,) = torch.autograd.grad(out, ut, grad_outputs=_t(p))
    return g.numpy()


def step_jacobian(scheme: po.Scheme, grid: po.OracleGrid, bc: object, dt: float, t: float, u: np.ndarray) -> np.ndarray:
    def fun(x: torch.Tensor) -> torch.Tensor:
        return ssprk33_advance(lambda tt, y: apply_operator(scheme, grid, bc, tt, y), float(dt), float(t), x)

    return torch.autograd.functional.jacobian(fun, _t(u)).numpy()


def adjoint_step(
    scheme: po.Scheme,
    grid: po.OracleGrid,
    bc: object,
    checkpoint: dict,
    p0: np.ndarray,
    *,
    maxit: int,
    apply_boundary_p: Callable[[float, np.ndarray, np.ndarray], np.ndarray] | None = None,
):
    """Generator of ``(m, t, dt, u, p)`` following timestepping.py:155-215."""
    chk = checkpoint[maxit]
    assert chk["m"] == maxit
    t = chk["t"]
    p = p0
    if apply_boundary_p is not None:
        p = apply_boundary_p(chk["t"], chk["u"], p)
    yield maxit, t, np.float64(0.0), chk["u"], p
    for m in range(maxit - 1, -1, -1):
        chk = checkpoint[m]
        dt = t - chk["t"]
        p = step_vjp(scheme, grid, bc, dt, chk["t"], chk["u"], p)
        if apply_boundary_p is not None:
            p = apply_boundary_p(chk["t"], chk["u"], p)
        t = chk["t"]
        yield m, t, dt, chk["u"], p
