"""GPU parity of the hand-derived discrete adjoint (through the C ABI).

Oracles: the torch-autograd twin of the NumPy oracle (oracle/torch_twin.py, itself pinned
to the reference's adjoint_step golden vectors), the golden vectors directly, and a
central finite-difference directional derivative (the reference's own precedent,
tests/test_finite_difference.py:153-159).

Tolerances: 1e-12 relative on smooth data.  On top-hat / rough data the derivative of the
WENO weights is conditioned like 1/(eps + beta) with eps = 1e-12, and two exact
differentiation methods of the same function already differ by ~1e-10 (see
tests/test_oracle_golden.py); those cases are held to 1e-9.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

import cases as C
from common import load_golden, max_rel, oracle_setup
from oracle import pyshocks_oracle as po
from oracle import torch_twin as tt
from test_gpu_kernels import dev, ghost_x, host, hotpath_for
from test_oracle_golden import ADJ, ADJ_KEYS, adjoint_setup

pytestmark = pytest.mark.gpu

# (round 2: ESWENO32 -- the reconstruction under every flux and the Burgers ESWENO32 scheme with its dissipative flux --
# has transposed kernels too, although the reference's adjoint drivers only ever run WENO-JS schemes)
CASES = C.rhs_cases()


def tol_for(case: C.Case) -> float:
    if case.rec == "esweno32":
        # eps of the bound ESWENO32 scheme is O(dx^2) and 1e-6 otherwise: the weights' derivative is well conditioned
        return 1.0e-11
    return 1.0e-12 if case.state == "smooth" else 1.0e-9


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.key)
def test_apply_operator_vjp_vs_autograd(case: C.Case) -> None:
    hp, scheme, grid, bc = hotpath_for(case, "fast")
    u = C.state_for(case)
    rng = np.random.default_rng(C.hash_key(case.key) % (2**32))
    v = rng.standard_normal(grid.nx)
    if case.bc == "dirichlet":
        hp.set_ghost(C.dirichlet_values(case, case.t, ghost_x(case, grid)))
    got = host(hp.apply_operator_vjp(dev(u), dev(v)))
    ref = tt.rhs_vjp(scheme, grid, bc, case.t, u, v)
    assert max_rel(got, ref) < tol_for(case), max_rel(got, ref)


@pytest.mark.parametrize(
    "case",
    [c for c in CASES if c.state == "smooth" and c.rec == "wenojs53"],
    ids=lambda c: c.key,
)
def test_apply_operator_vjp_vs_finite_differences(case: C.Case) -> None:
    """<v, J d> by central differences of the ORACLE forward == <J^T v, d> from the kernel"""
    hp, scheme, grid, bc = hotpath_for(case, "fast")
    u = C.state_for(case)
    rng = np.random.default_rng(5)
    v = rng.standard_normal(grid.nx)
    d = rng.standard_normal(grid.nx)
    if case.bc == "dirichlet":
        hp.set_ghost(C.dirichlet_values(case, case.t, ghost_x(case, grid)))
    jtv = host(hp.apply_operator_vjp(dev(u), dev(v)))
    h = 1.0e-6
    f = lambda x: po.apply_operator(scheme, grid, bc, case.t, x)  # noqa: E731
    jd = (f(u + h * d) - f(u - h * d)) / (2 * h)
    lhs, rhs = float(v @ jd), float(jtv @ d)
    assert abs(lhs - rhs) <= 1.0e-7 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)


@pytest.mark.parametrize("key", ADJ_KEYS)
def test_adjoint_step_sweep_vs_reference_golden(key: str) -> None:
    """The reference's adjoint_step (drivers/*-adjoint.py set-ups), every intermediate p."""
    from pyshocks_b200.path import HotPath

    scheme, grid, bc, pbc, theta, tfinal = adjoint_setup(key)
    maxit = int(ADJ[f"{key}_maxit"])
    chk_t, chk_u, p_ref = ADJ[f"{key}_chk_t"], ADJ[f"{key}_chk_u"], ADJ[f"{key}_p"]
    nu = grid.df ** (scheme.alpha - 1) if abs(scheme.alpha - 1.0) > 1.0e-8 else None
    hp = HotPath(equation=scheme.equation, flux=scheme.flux, rec=scheme.rec.name, bc=bc.kind,
                 n=grid.n, g=grid.g, dx=grid.h, eps=scheme.rec.eps, nu=nu, velocity=scheme.velocity)
    g = grid.g
    xg = np.concatenate([grid.x[:g], grid.x[grid.nx - g :]])

    def ghosts_at(t: float) -> np.ndarray | None:
        if isinstance(bc, po.Dirichlet):
            return np.concatenate([bc.ga(t, grid.x[:g]) + np.zeros(g), bc.ga(t, grid.x[grid.nx - g :]) + np.zeros(g)])
        return None

    p = po.apply_boundary(pbc, grid, chk_t[maxit], chk_u[maxit])  # p(T) = u(T), then BC on p
    assert np.array_equal(p, p_ref[0])
    t = chk_t[maxit]
    worst = 0.0
    for j, m in enumerate(range(maxit - 1, -1, -1), start=1):
        dt = t - chk_t[m]
        gh = None
        if isinstance(bc, po.Dirichlet):
            gh = [ghosts_at(tt_) for tt_ in (chk_t[m], chk_t[m] + dt, chk_t[m] + 0.5 * dt)]
        pd = hp.ssprk33_step_adjoint(dev(chk_u[m]), dev(np.array([dt])), dev(p), ghosts=gh)
        p = po.apply_boundary(pbc, grid, chk_t[m], host(pd))
        worst = max(worst, max_rel(p, p_ref[j]))
        t = chk_t[m]
    tol = 1.0e-9 if ("burgers" in key and "wenojs53" in key and "periodic" not in key) else 1.0e-12
    print(f"{key}: adjoint sweep of {maxit} steps, worst max-rel vs reference {worst:.3e}")
    assert worst < tol


@pytest.mark.parametrize("flux,rec", [("rusanov", "wenojs53"), ("esweno32", "esweno32"), ("godunov", "esweno32")])
def test_step_adjoint_is_transpose_of_step_jacobian(flux: str, rec: str) -> None:
    """dense check on a small periodic Burgers problem: (J^T p) for unit vectors p rebuilds J"""
    case = C.Case("burgers", flux, rec, "periodic", n=24, state="smooth")
    hp, scheme, grid, bc = hotpath_for(case, "fast")
    u = C.state_for(case)
    dt = 0.01
    J = tt.step_jacobian(scheme, grid, bc, dt, 0.0, u)
    P = np.eye(grid.nx)
    U = np.tile(u, (grid.nx, 1))
    out = host(hp.ssprk33_step_adjoint(dev(U), dev(np.array([dt])), dev(P)))
    # row r of out = J^T e_r = r-th row of J
    assert max_rel(out, J) < (1.0e-12 if rec == "wenojs53" else 1.0e-11)
