"""Pin the CPU oracles against golden vectors recorded from the reference's own code.

The golden files under tests/golden/ come from tests/golden/make_golden.py, which
executes /root/reference/src/pyshocks unmodified (over a NumPy-backed jax shim).
NumPy oracle: bit-for-bit.  C oracle: bit-for-bit on every cell whose stencil does
not reach past the array ends; the outermost two ghost rows per side involve
np.convolve's edge handling (a BLAS dot with its own summation order) and are held
to 1e-13 of the array maximum.  torch twin (adjoint): 1e-12 relative against the reference adjoint_step on smooth
data, 1e-9 on top-hat data (conditioning of the WENO weights, see the test).
"""

from __future__ import annotations

import numpy as np
import pytest

import cases as C
from common import load_golden, max_rel, oracle_setup
from oracle import pyshocks_oracle as po
from oracle import torch_twin as tt
from oracle.c_oracle import COracle

RHS = load_golden("rhs")
ADV = load_golden("advance")
CASES = C.rhs_cases()


def c_oracle_for(case: C.Case, scheme: po.Scheme, grid: po.OracleGrid, batch: int = 1) -> COracle:
    nu = grid.df ** (case.alpha - 1) if abs(case.alpha - 1.0) > 1.0e-8 else None
    return COracle(
        equation=case.equation, flux=case.flux, rec=case.rec, bc=case.bc, n=case.n, g=case.g,
        batch=batch, dx=grid.h, eps=scheme.rec.eps, delta=scheme.rec.delta, nu=nu, velocity=scheme.velocity,
    )


def ghost_x(case: C.Case, grid: po.OracleGrid) -> np.ndarray:
    return np.concatenate([grid.x[: case.g], grid.x[grid.nx - case.g :]])


def assert_c_matches(val: np.ndarray, ref: np.ndarray, edge: int) -> None:
    """bitwise away from the array ends, round-off level within `edge` entries of them"""
    core = slice(edge, val.size - edge)
    assert np.array_equal(val[core], ref[core])
    # ghost rows: round-off of np.convolve's edge dot products, amplified by 1/dx
    assert np.max(np.abs(val - ref)) <= 1.0e-13 * max(np.max(np.abs(ref)), 1.0)


def test_weno_pieces_bitwise() -> None:
    # weno.py:114-157, :247-256 and reconstruction.py:358-377
    G = load_golden("weno")
    for name, table in (("wenojs32", po._JS32), ("wenojs53", po._JS53)):
        rec = po.make_reconstruction(name)
        for label in ("sine", "step", "rough"):
            k = f"{name}_{label}"
            u = G[f"{k}_u"]
            assert np.array_equal(po.weno_smoothness(table, u), G[f"{k}_beta"])
            assert np.array_equal(po.weno_interp(table, u), G[f"{k}_uhat"])
            assert np.array_equal(po.weno_js_weights(table, u, rec.eps), G[f"{k}_omega"])
            ul, ur = po.reconstruct(rec, u)
            assert np.array_equal(ul, G[f"{k}_ul"])
            assert np.array_equal(ur, G[f"{k}_ur"])
    # ESWENO32: weno.py:284-296, reconstruction.py:413-439
    rec = po.make_reconstruction("esweno32")
    for label in ("sine", "step", "rough"):
        k = f"esweno32_{label}"
        u = G[f"{k}_u"]
        assert np.array_equal(po.es_weno_weights(po._JS32, u, rec.eps), G[f"{k}_omega"])
        ul, ur = po.reconstruct(rec, u)
        assert np.array_equal(ul, G[f"{k}_ul"])
        assert np.array_equal(ur, G[f"{k}_ur"])


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.key)
def test_rhs_numpy_oracle_bitwise(case: C.Case) -> None:
    scheme, grid, bc = oracle_setup(case)
    k = case.key
    u = RHS[f"{k}_u"]
    assert np.array_equal(u, C.state_for(case))
    if case.flux == "esweno32":  # bind took eps, delta from the grid (burgers/schemes.py:217-227)
        assert scheme.rec.eps == RHS[f"{k}_eps"] and scheme.rec.delta == RHS[f"{k}_delta"]
    w = po.apply_boundary(bc, grid, case.t, u)
    assert np.array_equal(w, RHS[f"{k}_w"])
    assert np.array_equal(po.numerical_flux(scheme, grid, w), RHS[f"{k}_f"])
    assert np.array_equal(po.apply_operator(scheme, grid, bc, case.t, u), RHS[f"{k}_L"])
    assert np.array_equal(po.predict_timestep(scheme, grid, u), RHS[f"{k}_dt"])
    if f"{k}_out" in ADV:
        dt = float(ADV[f"{k}_dt"])
        out = po.ssprk33_advance(lambda t, x: po.apply_operator(scheme, grid, bc, t, x), dt, case.t, u)
        assert np.array_equal(out, ADV[f"{k}_out"])


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.key)
def test_rhs_c_oracle(case: C.Case) -> None:
    scheme, grid, bc = oracle_setup(case)
    k = case.key
    u = RHS[f"{k}_u"]
    co = c_oracle_for(case, scheme, grid)
    xg = ghost_x(case, grid)
    if case.bc == "dirichlet":
        co.set_ghost(C.dirichlet_values(case, case.t, xg))
    w = co.apply_boundary(u)[0]
    assert np.array_equal(w, RHS[f"{k}_w"])
    assert_c_matches(co.numerical_flux(w)[0], RHS[f"{k}_f"], edge=3)
    assert_c_matches(co.apply_operator(u)[0], RHS[f"{k}_L"], edge=3)
    smax = co.max_abs(u, interior_only=True)[0]
    assert smax == np.max(np.abs(u[grid.interior]))
    if f"{k}_out" in ADV:
        dt = float(ADV[f"{k}_dt"])
        g3 = None
        if case.bc == "dirichlet":
            g3 = np.stack([C.dirichlet_values(case, tt_, xg) for tt_ in (case.t, case.t + dt, case.t + 0.5 * dt)])
        assert_c_matches(co.ssprk33_step(u, dt, g3)[0], ADV[f"{k}_out"], edge=3)


def test_c_oracle_batch_rows_independent() -> None:
    case = C.Case("burgers", "rusanov", "wenojs53", "periodic", state="rough")
    scheme, grid, bc = oracle_setup(case)
    rng = np.random.default_rng(3)
    U = np.stack([C.state_for(case) + 0.1 * rng.standard_normal(grid.nx) for _ in range(5)])
    co = c_oracle_for(case, scheme, grid, batch=5)
    L = co.apply_operator(U)
    for r in range(5):
        ref = po.apply_operator(scheme, grid, bc, 0.0, U[r])
        assert_c_matches(L[r], ref, edge=3)


@pytest.mark.parametrize("sname", ["rusanov", "lf"])
def test_config1_solve(sname: str) -> None:
    """BASELINE config 1: examples/burgers.py, WENOJS53, periodic, N=256, t=1 (171 steps)."""
    S = load_golden("solve_c1")
    grid = po.make_grid(-1.5, 1.5, 256, 3)
    rec = po.make_reconstruction("wenojs53")
    scheme = po.Scheme("burgers", sname, rec)
    bc = po.Periodic()
    u0 = po.cell_average(grid, lambda x: po.burgers_tophat(grid, 0.0, x), 4)
    assert np.array_equal(u0, S[f"{sname}_u0"])

    # NumPy oracle: the whole trajectory is bit-identical to the reference's
    dts = []
    snaps = {}
    for m, t, dt, u in po.step(
        lambda t_, u_: po.apply_operator(scheme, grid, bc, t_, u_),
        lambda t_, u_: 1.0 * po.predict_timestep(scheme, grid, u_),
        u0, tfinal=1.0,
    ):
        dts.append(dt)
        snaps[m] = u
    assert m == 171
    assert np.array_equal(np.array(dts), S[f"{sname}_dt"])
    for it in (1, 10, 50, 100):
        assert np.array_equal(snaps[it], S[f"{sname}_u{it:03d}"])
    assert np.array_equal(u, S[f"{sname}_uf"])

    # C oracle: same dt history, interior bit-identical
    co = COracle(equation="burgers", flux=sname, rec="wenojs53", bc="periodic", n=256, g=3,
                 batch=1, dx=grid.h, eps=rec.eps)
    uf, hist = co.solve_adaptive(u0, 1.0, 0.5 * grid.h ** (2 - 1.0), 1.0)
    assert np.array_equal(hist, S[f"{sname}_dt"][1:])
    assert np.array_equal(uf[grid.interior], S[f"{sname}_uf"][grid.interior])


def test_config2_forward_reduced() -> None:
    """drivers/advection-adjoint.py forward set-up at N=128."""
    S = load_golden("solve_c2")
    grid = po.make_grid(-1.0, 1.0, 128, 3)
    rec = po.make_reconstruction("wenojs53")
    vel = po.cell_average(grid, lambda x: np.full_like(x, 1.0), 4)
    assert np.array_equal(vel, S["velocity"])
    scheme = po.Scheme("advection", "godunov", rec, velocity=vel)
    bc = po.Dirichlet(ga=lambda t, x: po.ic_sine(grid, x - 1.0 * t, k=1))
    u0 = po.cell_average(grid, lambda x: po.ic_sine(grid, x, k=1), 4)
    assert np.array_equal(u0, S["u0"])
    dts = []
    for m, t, dt, u in po.step(
        lambda t_, u_: po.apply_operator(scheme, grid, bc, t_, u_),
        lambda t_, u_: 0.75 * po.predict_timestep(scheme, grid, u_),
        u0, tfinal=1.0,
    ):
        dts.append(dt)
    assert np.array_equal(np.array(dts), S["dt"])
    assert np.array_equal(u, S["uf"])


ADJ = load_golden("adjoint")
ADJ_KEYS = sorted({k.rsplit("_u0", 1)[0] for k in ADJ if k.endswith("_u0")})


def adjoint_setup(key: str):
    parts = key.split("_")
    eq, flux, recn = parts[0], parts[1], parts[2]
    alpha = 0.995 if "alpha" in key else 1.0
    periodic = key.endswith("periodic")
    rec = po.make_reconstruction(recn)
    if eq == "burgers":
        grid = po.make_grid(-1.5, 1.5, 48, rec.stencil_width)
        scheme = po.Scheme("burgers", flux, rec, alpha=alpha)
        bc = po.Periodic() if periodic else po.Dirichlet(ga=lambda t, x: po.burgers_tophat(grid, t, x))
        pbc = po.Periodic() if periodic else po.Neumann(ga=lambda t: 0.0)
        theta, tfinal = 1.0, 0.4
    else:
        grid = po.make_grid(-1.0, 1.0, 48, 3)
        vel = po.cell_average(grid, lambda x: np.ones_like(x), 4)
        scheme = po.Scheme("advection", "godunov", rec, velocity=vel)
        bc = po.Periodic() if periodic else po.Dirichlet(ga=lambda t, x: po.ic_sine(grid, x - 1.0 * t, k=1))
        pbc = po.Periodic() if periodic else po.Dirichlet(ga=lambda t, x: np.zeros_like(x))
        theta, tfinal = 0.75, 0.5
    return scheme, grid, bc, pbc, theta, tfinal


@pytest.mark.parametrize("key", ADJ_KEYS)
def test_adjoint_step_torch_twin_vs_reference(key: str) -> None:
    """timestepping.adjoint_step (jacfwd + J^T p) of the reference vs autograd of the twin."""
    scheme, grid, bc, pbc, theta, tfinal = adjoint_setup(key)
    maxit = int(ADJ[f"{key}_maxit"])
    chk: dict = {}
    for m, t, dt, u in po.step(
        lambda t_, u_: po.apply_operator(scheme, grid, bc, t_, u_),
        lambda t_, u_: theta * po.predict_timestep(scheme, grid, u_),
        ADJ[f"{key}_u0"], tfinal=tfinal, checkpoint=chk,
    ):
        pass
    assert m == maxit
    assert np.array_equal(np.stack([chk[i]["u"] for i in range(maxit + 1)]), ADJ[f"{key}_chk_u"])
    assert np.array_equal(np.array([chk[i]["t"] for i in range(maxit + 1)]), ADJ[f"{key}_chk_t"])
    ps = np.stack([
        p for (_, _, _, _, p) in tt.adjoint_step(
            scheme, grid, bc, chk, chk[maxit]["u"], maxit=maxit,
            apply_boundary_p=lambda t, u, p: po.apply_boundary(pbc, grid, t, p),
        )
    ])
    # Smooth data: round-off agreement.  Top-hat data (flat regions next to a jump,
    # eps = 1e-12): d(omega)/d(beta) ~ 1/(eps + beta) amplifies the 1e-16 round-off of
    # (q_k - R) to ~1e-10, so complex-step (golden) and reverse-mode (twin) evaluations
    # of the SAME derivative differ at that level ("adjoint conditioning").
    tol = 1.0e-9 if ("burgers" in key and "wenojs53" in key and "periodic" not in key) else 1.0e-12
    assert max_rel(ps, ADJ[f"{key}_p"]) < tol


@pytest.mark.parametrize("case", C.stepper_cases(), ids=lambda c: c.key)
def test_other_steppers_numpy_oracle_bitwise(case: C.Case) -> None:
    """ForwardEuler / RK44 / CKRK45 ``advance`` (timestepping.py:289-405) of the NumPy restatement against
    the vectors recorded from the reference (steppers.npz); the Dirichlet cases cover the stage times."""
    G = load_golden("steppers")
    scheme, grid, bc = oracle_setup(case)
    k = case.key
    u, dt = G[f"{k}_u"], float(G[f"{k}_dt"])
    assert np.array_equal(u, C.state_for(case))
    for name in C.STEPPERS:
        out = po.STEPPER_ADVANCE[name](lambda t, x: po.apply_operator(scheme, grid, bc, t, x), dt, case.t, u)
        assert np.array_equal(out, G[f"{k}_{name}"]), name


@pytest.mark.parametrize("case", [c for c in C.rhs_cases() if c.rec == "esweno32"], ids=lambda c: c.key)
def test_torch_twin_esweno32_is_the_numpy_oracle(case: C.Case) -> None:
    """the autograd twin of the ESWENO32 reconstruction and of the Burgers ESWENO32 scheme (weno.py:284-296,
    burgers/schemes.py:230-256) evaluates the NumPy oracle's RHS -- itself pinned to the reference's golden vectors --
    to the last bit: its reverse-mode derivative is the truth the transposed ESWENO32 kernels are held to"""
    import torch

    from common import oracle_setup
    from oracle import torch_twin as tt

    scheme, grid, bc = oracle_setup(case)
    u = C.state_for(case)
    L = po.apply_operator(scheme, grid, bc, case.t, u)
    Lt = tt.apply_operator(scheme, grid, bc, case.t, torch.from_numpy(u)).numpy()
    assert np.array_equal(L, Lt)
    # and the vector-Jacobian product agrees with central differences of the oracle
    rng = np.random.default_rng(1)
    v, d = rng.standard_normal(grid.nx), rng.standard_normal(grid.nx)
    jtv = tt.rhs_vjp(scheme, grid, bc, case.t, u, v)
    if case.state == "smooth":
        h = 1e-6
        f = lambda x: po.apply_operator(scheme, grid, bc, case.t, x)  # noqa: E731
        jd = (f(u + h * d) - f(u - h * d)) / (2 * h)
        lhs, rhs = float(v @ jd), float(jtv @ d)
        assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs), 1.0)
