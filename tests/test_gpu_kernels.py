"""GPU parity of the forward kernels (through the C ABI) against the golden vectors
recorded from the reference and against the CPU oracles.

STRICT math: bit-for-bit on every cell whose stencil does not touch the zero padding at
the array ends (the same criterion the C oracle is held to); round-off level on the
outermost ghost rows.  FAST math: tolerance stated per test.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

import cases as C
from common import load_golden, max_rel, oracle_setup
from oracle import pyshocks_oracle as po

pytestmark = pytest.mark.gpu

RHS = load_golden("rhs")
ADV = load_golden("advance")
CASES = C.rhs_cases()

# FAST vs reference, one RHS evaluation, relative to max |L|: re-association + FMA +
# 1-ulp reciprocal; measured <= 3e-14 on these cases
FAST_RHS_TOL = 5.0e-13


def hotpath_for(case: C.Case, math: str):
    from pyshocks_b200.path import HotPath

    scheme, grid, bc = oracle_setup(case)
    nu = grid.df ** (case.alpha - 1) if abs(case.alpha - 1.0) > 1.0e-8 else None
    hp = HotPath(
        equation=case.equation, flux=case.flux, rec=case.rec, bc=case.bc, n=case.n, g=case.g,
        dx=grid.h, eps=scheme.rec.eps, math=math, nu=nu, velocity=scheme.velocity, delta=scheme.rec.delta,
    )
    return hp, scheme, grid, bc


def ghost_x(case: C.Case, grid: po.OracleGrid) -> np.ndarray:
    return np.concatenate([grid.x[: case.g], grid.x[grid.nx - case.g :]])


def dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def assert_strict(val: np.ndarray, ref: np.ndarray, edge: int = 3) -> None:
    core = slice(edge, val.size - edge)
    assert np.array_equal(val[core], ref[core])
    assert np.max(np.abs(val - ref)) <= 1.0e-13 * max(np.max(np.abs(ref)), 1.0)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.key)
def test_rhs_strict_matches_reference_bitwise(case: C.Case) -> None:
    hp, scheme, grid, bc = hotpath_for(case, "strict")
    k = case.key
    u = dev(RHS[f"{k}_u"])
    if case.bc == "dirichlet":
        hp.set_ghost(C.dirichlet_values(case, case.t, ghost_x(case, grid)))
    w = hp.apply_boundary(u)
    assert np.array_equal(host(w), RHS[f"{k}_w"])
    assert_strict(host(hp.numerical_flux(w)), RHS[f"{k}_f"])
    assert_strict(host(hp.apply_operator(u)), RHS[f"{k}_L"])
    smax = host(hp.max_abs(u, 1))[0]
    assert smax == np.max(np.abs(RHS[f"{k}_u"][grid.interior]))


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.key)
def test_rhs_fast_within_tolerance(case: C.Case) -> None:
    hp, scheme, grid, bc = hotpath_for(case, "fast")
    k = case.key
    u = dev(RHS[f"{k}_u"])
    if case.bc == "dirichlet":
        hp.set_ghost(C.dirichlet_values(case, case.t, ghost_x(case, grid)))
    L = host(hp.apply_operator(u))
    assert max_rel(L, RHS[f"{k}_L"]) < FAST_RHS_TOL
    F = host(hp.numerical_flux(hp.apply_boundary(u)))
    assert max_rel(F, RHS[f"{k}_f"]) < FAST_RHS_TOL


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: c.key)
def test_rhs_axpby_is_the_combine_of_the_reference_rhs(case: C.Case, math: str) -> None:
    """psk_rhs_axpby: ``ca u0 + cb u + cc dt L(u)`` in one launch, the building block of the fused RK44 / CKRK45
    stages (timestepping.py:289-405), against the same expression over the recorded reference ``L``: bitwise in
    STRICT math (every operation rounded once, left to right), within the RHS tolerance in FAST math."""
    hp, scheme, grid, bc = hotpath_for(case, math)
    k = case.key
    u_h = RHS[f"{k}_u"]
    rng = np.random.default_rng(17)
    u0_h = u_h + 0.1 * rng.standard_normal(u_h.shape)
    if case.bc == "dirichlet":
        hp.set_ghost(C.dirichlet_values(case, case.t, ghost_x(case, grid)))
    dt = 0.37 * grid.h
    for ca, cb, cc in [(1.0, 0.0, 0.5), (-0.4178904745, 0.0, 1.0), (1.0, 1.0 / 3.0, 1.0 / 6.0), (0.0, 1.0, 1.0)]:
        ref = ca * u0_h + cb * u_h + cc * (dt * RHS[f"{k}_L"])
        out = host(hp.rhs_axpby(dev(u0_h), dev(u_h), dev(np.array([dt])), ca, cb, cc, out=torch.zeros_like(dev(u_h))))
        it = grid.interior
        if math == "strict":
            assert_strict(out[it], ref[it])
        else:
            scale = max(np.max(np.abs(dt * RHS[f"{k}_L"])), 1.0e-300)
            assert np.max(np.abs(out[it] - ref[it])) < FAST_RHS_TOL * max(scale, np.max(np.abs(ref[it])))
    # the output must not alias the RHS input: its halo is read while neighbours are written
    uu = dev(u_h)
    with pytest.raises(Exception, match="psk_rhs_axpby"):
        hp.rhs_axpby(dev(u0_h), uu, dev(np.array([dt])), 1.0, 0.0, 1.0, out=uu)


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("case", [c for c in CASES if f"{c.key}_out" in ADV], ids=lambda c: c.key)
def test_ssprk33_step_matches_reference_advance(case: C.Case, math: str) -> None:
    hp, scheme, grid, bc = hotpath_for(case, math)
    k = case.key
    u = dev(ADV[f"{k}_u"])
    dt = float(ADV[f"{k}_dt"])
    ghosts = None
    if case.bc == "dirichlet":
        xg = ghost_x(case, grid)
        ghosts = [C.dirichlet_values(case, tt, xg) for tt in (case.t, case.t + dt, case.t + 0.5 * dt)]
    out = host(hp.ssprk33_step(u, dev(np.array([dt])), ghosts=ghosts, ghost_rows=True))
    ref = ADV[f"{k}_out"]
    if math == "strict":
        assert_strict(out, ref)
    else:
        assert max_rel(out, ref) < 1.0e-13


def test_reconstruct_matches_reference() -> None:
    from pyshocks_b200.path import HotPath

    G = load_golden("weno")
    for name, g in (("wenojs32", 2), ("wenojs53", 3), ("esweno32", 2)):
        eps = po.make_reconstruction(name).eps
        for label in ("sine", "step", "rough"):
            k = f"{name}_{label}"
            u = G[f"{k}_u"]
            for math in ("strict", "fast"):
                hp = HotPath(equation="burgers", flux="rusanov", rec=name, bc="none", n=u.size - 2 * g,
                             g=g, dx=1.0, eps=eps, math=math)
                ul, ur = hp.reconstruct(dev(u))
                if math == "strict":
                    assert_strict(host(ul), G[f"{k}_ul"], edge=2)
                    assert_strict(host(ur), G[f"{k}_ur"], edge=2)
                else:
                    assert max_rel(host(ul), G[f"{k}_ul"]) < 1.0e-14
                    assert max_rel(host(ur), G[f"{k}_ur"]) < 1.0e-14


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("ld_pad", [0, 1, 10])
def test_batched_rows_equal_single_rows(math: str, ld_pad: int) -> None:
    """ensemble rows are independent reference runs; any row stride / alignment works"""
    case = C.Case("burgers", "rusanov", "wenojs53", "periodic", n=1100, state="rough")
    hp, scheme, grid, bc = hotpath_for(case, math)
    rng = np.random.default_rng(7)
    B = 5
    U = np.stack([C.state_for(case) + 0.2 * rng.standard_normal(grid.nx) for _ in range(B)])
    store = torch.zeros((B, grid.nx + ld_pad), dtype=torch.float64, device="cuda")
    Ud = store[:, : grid.nx]
    Ud.copy_(dev(U))
    dt = dev(np.linspace(1e-3, 2e-3, B))
    out = torch.zeros_like(store)[:, : grid.nx]
    hp.ssprk33_step(Ud, dt, out=out, ghost_rows=True)
    for r in range(B):
        single = hp.ssprk33_step(dev(U[r]), dt[r : r + 1].clone(), ghost_rows=True)
        assert torch.equal(out[r], single)
    if math == "strict":
        from oracle.c_oracle import COracle

        co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=case.n, g=3,
                     batch=B, dx=grid.h, eps=scheme.rec.eps)
        ref = co.ssprk33_step(U, host(dt))
        assert np.array_equal(host(out), ref)


def test_inactive_rows_untouched_and_fused_cfl_max() -> None:
    case = C.Case("burgers", "rusanov", "wenojs53", "periodic", n=300, state="smooth")
    hp, scheme, grid, bc = hotpath_for(case, "fast")
    rng = np.random.default_rng(11)
    B = 6
    U = dev(np.stack([C.state_for(case) * (1 + r) for r in range(B)]))
    active = torch.tensor([1, 0, 1, 1, 0, 1], dtype=torch.uint8, device="cuda")
    maxabs = torch.zeros(B, dtype=torch.float64, device="cuda")
    dt = dev(np.full(B, 1e-3))
    out = hp.ssprk33_step(U, dt, active=active, maxabs=maxabs)
    for r in range(B):
        if active[r] == 0:
            assert torch.equal(out[r], U[r])
            assert maxabs[r] == 0.0
        else:
            assert maxabs[r] == out[r, grid.interior].abs().max()


def test_step_control_matches_reference_loop() -> None:
    """timestepping.py:139-150 per row, against the NumPy statement"""
    import ctypes as ct

    from pyshocks_b200 import _lib as L

    B = 7
    maxabs = np.array([1.0, 0.5, 2.0, 0.0, np.inf, 1e-3, 3.0])
    t = np.array([0.0, 0.99999, 0.5, 0.2, 0.3, 0.1, 1.0])
    theta, scale, tfinal = 0.9, 0.5 * (3.0 / 256), 1.0
    md, td = dev(maxabs), dev(t)
    tn, dt = torch.zeros_like(td), torch.zeros_like(td)
    act = torch.zeros(B, dtype=torch.uint8, device="cuda")
    nonf = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.check("psk_step_control", L.lib().psk_step_control(
        B, theta, scale, tfinal, L.ptr(md), L.ptr(td), L.ptr(tn), L.ptr(dt), L.raw_ptr(act), L.raw_ptr(nonf),
        L.stream_ptr()))
    with np.errstate(divide="ignore"):
        for r in range(B):
            if t[r] >= tfinal:
                assert act[r] == 0 and tn[r] == t[r]
                continue
            d = theta * (scale / maxabs[r])
            d = (d if d < tfinal - t[r] else tfinal - t[r]) + 1.0e-15
            assert float(dt[r]) == d and float(tn[r]) == t[r] + d and act[r] == 1
    assert int(nonf[0]) == 0


@pytest.mark.parametrize("sname", ["rusanov", "lf"])
def test_config1_strict_solve_is_bitwise_reference(sname: str) -> None:
    """BASELINE config 1 (examples/burgers.py, N=256, t=1): 171 adaptive steps driven by the
    fused CFL reduction + device-side step control reproduce the reference bit for bit."""
    from pyshocks_b200.ensemble import EnsembleSolver

    S = load_golden("solve_c1")
    grid = po.make_grid(-1.5, 1.5, 256, 3)
    for math in ("strict", "fast"):
        solver = EnsembleSolver(equation="burgers", flux=sname, rec="wenojs53", bc="periodic",
                                n=256, g=3, dx=grid.h, eps=1.0e-12, math=math, batch=1)
        res = solver.solve_adaptive(dev(S[f"{sname}_u0"][None, :]), theta=1.0, tfinal=1.0,
                                    cfl_scale=0.5 * grid.h ** (2 - 1.0), record_dt=True)
        uf = host(res.u)[0]
        if math == "strict":
            assert res.steps == 171
            assert np.array_equal(res.dt_history[:, 0], S[f"{sname}_dt"][1:])
            assert np.array_equal(uf[grid.interior], S[f"{sname}_uf"][grid.interior])
        else:
            assert res.steps == 171
            err = max_rel(uf[grid.interior], S[f"{sname}_uf"][grid.interior])
            print(f"config-1 {sname} FAST vs reference after 171 steps: max rel {err:.3e}")
            assert err < 1.0e-12


def test_host_to_host_pipelined_call_equals_device_resident_solve() -> None:
    """the e2e entry point (row blocks on several streams, copies overlapped with compute) gives
    the same bits as load -> solve -> store"""
    from pyshocks_b200.ensemble import EnsembleSolver

    B, n, g, nsteps = 37, 1024, 3, 6
    rng = np.random.default_rng(5)
    u0 = 0.5 + 0.4 * rng.standard_normal((B, n + 2 * g))
    kw = dict(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=3.0 / n, eps=1e-12, batch=B)
    ref = EnsembleSolver(**kw).solve_fixed_dt(dev(u0), 2e-4, nsteps).u.cpu()
    solver = EnsembleSolver(**kw)
    hin = torch.from_numpy(u0).pin_memory()
    hout = torch.empty_like(hin).pin_memory()
    solver.solve_fixed_dt_host(hin, hout, 2e-4, nsteps, groups=5, streams=3)
    torch.cuda.synchronize()
    i = slice(g, g + n)
    assert torch.equal(hout[:, i], ref[:, i])


@pytest.mark.parametrize("name,g", [("wenojs32", 2), ("wenojs53", 3)])
def test_kernel_coefficients_are_those_of_the_stencil_tables(name: str, g: int) -> None:
    """The WENO-JS coefficients compiled into the kernels (csrc/psk_math.cuh) against the tables the package
    exposes (pyshocks_b200/weno.py = pyshocks/weno.py:166-244), evaluated the way the reference evaluates them
    (weno.py:114-157, :247-256: zero-padded "same" convolutions with the table rows): STRICT reconstruction
    bit for bit away from the array ends, FAST to round-off."""
    from pyshocks_b200 import weno
    from pyshocks_b200.path import HotPath
    from pyshocks_b200.reconstruction import make_reconstruction_from_name

    rec = make_reconstruction_from_name(name)
    s = weno.weno_js_32_coefficients() if name == "wenojs32" else weno.weno_js_53_coefficients()
    assert rec.s.c.shape == s.c.shape and np.array_equal(rec.s.d, s.d)
    rng = np.random.default_rng(53)
    x = np.linspace(0.0, 1.0, 97)
    u = np.sin(2 * np.pi * x) + 0.3 * rng.standard_normal(x.size) * (x > 0.6)

    def side(f: np.ndarray) -> np.ndarray:
        beta = [sum(s.a[j] * np.convolve(f, s.b[i, j, :], mode="same") ** 2 for j in range(s.a.size))
                for i in range(s.b.shape[0])]
        alpha = [s.d[i, 0] / (rec.eps + beta[i]) ** 2 for i in range(len(beta))]
        total = sum(alpha[1:], alpha[0])
        return sum(((alpha[i] / total) * np.convolve(f, s.c[i, :], mode="same") for i in range(1, len(beta))),
                   (alpha[0] / total) * np.convolve(f, s.c[0, :], mode="same"))

    ur = side(u)
    ul = side(u[::-1])[::-1]  # reconstruction.py:374-375
    for math, tol in (("strict", 0.0), ("fast", 1.0e-14)):
        hp = HotPath(equation="burgers", flux="rusanov", rec=name, bc="none", n=u.size - 2 * g, g=g, dx=1.0,
                     eps=rec.eps, math=math)
        gl, gr = (host(a) for a in hp.reconstruct(dev(u)))
        i = slice(2, u.size - 2)  # the outermost cells see NumPy's BLAS summation order
        if tol == 0.0:
            assert np.array_equal(gl[i], ul[i]) and np.array_equal(gr[i], ur[i])
        else:
            assert max_rel(gl[i], ul[i]) < tol and max_rel(gr[i], ur[i]) < tol
