"""The pyshocks-shaped API on the GPU: the reference's scripts and tests, re-run through
``pyshocks_b200`` (same names, same call sequences), checked against golden vectors recorded
from the reference and against the reference tests' own acceptance criteria."""

from __future__ import annotations

from functools import partial

import numpy as np
import pytest
import torch

from common import load_golden, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fast_math():
    from pyshocks_b200 import config

    config.set_math("fast")
    yield
    config.set_math("fast")


def host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


# {{{ examples/burgers.py (BASELINE config 1)


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("sname", ["rusanov", "lf"])
def test_examples_burgers_config1(sname: str, math: str) -> None:
    """examples/burgers.py:42-60,120-193 with -s rusanov|lf -r wenojs53 -n 256, tfinal=1."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import burgers, config, funcs, timestepping
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import PeriodicBoundary

    config.set_math(math)
    S = load_golden("solve_c1")
    rec = make_reconstruction_from_name("wenojs53")
    scheme = burgers.make_scheme_from_name(sname, rec=rec, alpha=1.0)
    order = int(max(scheme.order, 1.0)) + 1
    grid = ps.make_uniform_cell_grid(a=-1.5, b=1.5, n=256, nghosts=scheme.stencil_width)
    quad = ps.make_leggauss_quadrature(grid, order=order)
    u0 = ps.cell_average(quad, lambda x: funcs.burgers_tophat(grid, 0.0, x))
    assert np.array_equal(host(u0), S[f"{sname}_u0"])
    boundary = PeriodicBoundary()
    scheme = ps.bind(scheme, grid, boundary)
    theta = 1.0

    def _predict_timestep(t_, u_):
        return theta * ps.predict_timestep(scheme, grid, boundary, t_, u_)

    def _apply_operator(t_, u_):
        return ps.apply_operator(scheme, grid, boundary, t_, u_)

    method = timestepping.SSPRK33(
        predict_timestep=ps.jit(_predict_timestep), source=ps.jit(_apply_operator), checkpoint=None
    )
    dts = []
    for event in timestepping.step(method, u0, tfinal=1.0):
        dts.append(float(event.dt))
        energy = ps.norm(grid, event.u, p=2, weighted=True)
        tv = ps.norm(grid, event.u, p="tvd")
    assert event.iteration == 171
    assert method.source.bound is not None  # the jit() wrapper recognised apply_operator -> fused stages
    uf = host(event.u)
    i = grid.i_
    if math == "strict":
        assert np.array_equal(np.array(dts), S[f"{sname}_dt"])
        assert np.array_equal(uf[i], S[f"{sname}_uf"][i])
        assert max_rel(uf, S[f"{sname}_uf"]) < 1e-13  # ghost rows included
    else:
        assert max_rel(np.array(dts), S[f"{sname}_dt"]) < 1e-12
        assert max_rel(uf[i], S[f"{sname}_uf"][i]) < 1e-12
    assert float(energy) > 0 and float(tv) > 0


def test_fused_advance_equals_generic_advance() -> None:
    """advance() through an opaque callable (3 RHS launches + torch axpys, the reference's own
    structure) and through the fused stage kernels agree bit for bit in STRICT mode."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import burgers, config, funcs, timestepping
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import make_dirichlet_boundary

    config.set_math("strict")
    rec = make_reconstruction_from_name("wenojs53")
    scheme = burgers.make_scheme_from_name("rusanov", rec=rec)
    grid = ps.make_uniform_cell_grid(a=-1.5, b=1.5, n=200, nghosts=3)
    bc = make_dirichlet_boundary(ga=lambda t, x: funcs.burgers_tophat(grid, t, x))
    quad = ps.make_leggauss_quadrature(grid, order=4)
    u0 = ps.cell_average(quad, lambda x: funcs.burgers_tophat(grid, 0.0, x))
    opaque = timestepping.SSPRK33(
        predict_timestep=lambda t, u: 1e-3,
        source=lambda t, u: ps.apply_operator(scheme, grid, bc, t, u) + 0.0,  # "+ 0.0": not recognisable
        checkpoint=None,
    )
    fused = timestepping.SSPRK33(
        predict_timestep=lambda t, u: 1e-3, source=ps.bind_operator(scheme, grid, bc), checkpoint=None
    )
    dt = 2.0e-3
    a = timestepping.advance(opaque, dt, 0.1, u0)
    b = timestepping.advance(fused, dt, 0.1, u0)
    assert torch.equal(a, b)


# }}}

# {{{ drivers/*-adjoint.py


@pytest.mark.parametrize("which", ["burgers", "advection"])
def test_adjoint_drivers_vs_reference(which: str) -> None:
    """drivers/burgers-adjoint.py:68-97,205-212,269-315 and drivers/advection-adjoint.py:219-328 at
    N=48: forward with an InMemoryCheckpoint, then adjoint_step with the drivers' BC on p."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import advection, burgers, funcs, timestepping
    from pyshocks_b200.checkpointing import InMemoryCheckpoint
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import make_dirichlet_boundary, make_neumann_boundary

    A = load_golden("adjoint")
    rec = make_reconstruction_from_name("wenojs53")
    if which == "burgers":
        key, tfinal, theta, tol = "burgers_rusanov_wenojs53", 0.4, 1.0, 1e-9
        scheme = burgers.make_scheme_from_name("rusanov", rec=rec, alpha=1.0)
        grid = ps.make_uniform_cell_grid(a=-1.5, b=1.5, n=48, nghosts=scheme.stencil_width)
        quad = ps.make_leggauss_quadrature(grid, order=int(max(scheme.order, 1)) + 1)
        u0 = ps.cell_average(quad, lambda x: funcs.burgers_tophat(grid, 0.0, x))
        bc = make_dirichlet_boundary(ga=lambda t, x: funcs.burgers_tophat(grid, t, x))
        pbc = make_neumann_boundary(lambda t: 0.0)
    else:
        key, tfinal, theta, tol = "advection_godunov_wenojs53_dirichlet", 0.5, 0.75, 1e-12
        scheme = advection.make_scheme_from_name("godunov", rec=rec, velocity=None)
        grid = ps.make_uniform_cell_grid(a=-1.0, b=1.0, n=48, nghosts=scheme.stencil_width)
        quad = ps.make_leggauss_quadrature(grid, order=int(max(scheme.order, 1.0)) + 1)
        from dataclasses import replace

        velocity = ps.cell_average(quad, partial(funcs.ic_constant, grid, c=1.0))
        scheme = replace(scheme, velocity=velocity)
        func_ic = partial(funcs.ic_sine, grid, k=1)
        u0 = ps.cell_average(quad, func_ic)
        bc = make_dirichlet_boundary(lambda t, x: func_ic(x - 1.0 * t))
        pbc = make_dirichlet_boundary(lambda t, x: torch.zeros_like(x))
    # initial data are not on the hot path: torch's device sin() may differ from libm in the
    # last bit, so start from the reference's own u0 (the Dirichlet data keep the device sin)
    assert max_rel(host(u0), A[f"{key}_u0"]) < 1e-15
    u0 = torch.from_numpy(A[f"{key}_u0"]).cuda()

    stepper = timestepping.SSPRK33(
        predict_timestep=ps.jit(lambda t_, u_: theta * ps.predict_timestep(scheme, grid, bc, t_, u_)),
        source=ps.jit(lambda t_, u_: ps.apply_operator(scheme, grid, bc, t_, u_)),
        checkpoint=InMemoryCheckpoint(basename="Iteration"),
    )
    for event in timestepping.step(stepper, u0, tfinal=tfinal):
        pass
    uf, maxit = event.u, event.iteration
    assert maxit == int(A[f"{key}_maxit"])
    assert max_rel(host(uf), A[f"{key}_chk_u"][maxit]) < 1e-12
    ps_ = []
    for ev in timestepping.adjoint_step(
        stepper, uf, maxit=maxit, apply_boundary=lambda t, u, p: ps.apply_boundary(pbc, grid, t, p)
    ):
        ps_.append(host(ev.p))
    err = max_rel(np.stack(ps_), A[f"{key}_p"])
    print(f"{which}-adjoint driver: {maxit} steps, p vs reference max rel {err:.3e}")
    assert err < tol


# }}}

# {{{ the reference's own tests, verbatim criteria


@pytest.mark.parametrize("rec_name", ["constant", "wenojs32", "wenojs53"])
@pytest.mark.parametrize("bc_type", ["periodic", "dirichlet"])
def test_advection_vs_continuity(rec_name: str, bc_type: str) -> None:
    """tests/test_finite_difference.py:30-84: |<u, A v> - <C u, v>| < 1e-15 for constant velocity."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import advection, config, continuity
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import PeriodicBoundary, make_dirichlet_boundary

    config.set_math("strict")
    rec = make_reconstruction_from_name(rec_name)
    grid = ps.make_uniform_cell_grid(a=-1.0, b=1.0, n=256, nghosts=rec.stencil_width)
    boundary = PeriodicBoundary() if bc_type == "periodic" else make_dirichlet_boundary(lambda t, x: torch.zeros_like(x))
    velocity = torch.ones_like(grid.x)
    ascheme = advection.Godunov(rec=rec, velocity=velocity)
    cscheme = continuity.Godunov(rec=rec, velocity=velocity)
    i = grid.i_

    def dot(x, y):
        return (x[i] * grid.dx[i]) @ y[i]

    aop = partial(ps.apply_operator, ascheme, grid, boundary, 0.0)
    cop = partial(ps.apply_operator, cscheme, grid, boundary, 0.0)
    u = torch.sin(2.0 * np.pi * grid.x)
    v = torch.sin(2.0 * np.pi * grid.x)
    error = abs(float(dot(u, aop(v)) - dot(cop(u), v)))
    assert error < 1.0e-15


def _evolve(case_scheme, case_bc, case_exact, n, *, dt, a=-1.0, b=1.0, tfinal=0.5):
    """tests/test_convergence.py:111-214 (the `evolve` driver) without the plotting."""
    import pyshocks_b200 as ps
    import pyshocks_b200.timestepping as ts

    grid = ps.make_uniform_cell_grid(a=a, b=b, n=n, nghosts=3)
    bc = case_bc(grid)
    scheme = ps.bind(case_scheme(grid, bc), grid, bc)
    quad = ps.make_leggauss_quadrature(grid, order=5)
    u0 = ps.cell_average(quad, lambda x: case_exact(grid, 0.0, x))
    maxit, dt = ts.predict_maxit_from_timestep(tfinal, dt)
    stepper = ts.SSPRK33(
        predict_timestep=lambda _t, _u: dt,
        source=ps.jit(lambda t_, u_: ps.apply_operator(scheme, grid, bc, t_, u_)),
        checkpoint=None,
    )
    u = u0
    for event in ts.step(stepper, u0, maxit=maxit):
        u = event.u
    uhat = ps.cell_average(quad, lambda x: case_exact(grid, tfinal, x))
    h_max = float(torch.max(torch.diff(grid.f)))
    error = float(ps.norm(grid, u - uhat, weighted=True) / ps.norm(grid, uhat, weighted=True))
    return h_max, error


@pytest.mark.parametrize(
    ("rec_name", "order", "resolutions"),
    [
        ("constant", 1, list(range(80, 160 + 1, 16))),
        ("wenojs32", 3, list(range(192, 384 + 1, 32))),
        ("wenojs53", 5, list(range(32, 256 + 1, 32))),
        ("esweno32", 3, list(range(32, 256 + 1, 32))),  # tests/test_convergence.py:402
    ],
)
def test_advection_convergence(rec_name: str, order: int, resolutions: list[int]) -> None:
    """tests/test_convergence.py:312-343,395-460: godunov + rec, periodic, ic_sine_sine, t=1,
    dt = 8 (2/n)^(5/3): EOC >= order - 0.5."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import advection, funcs
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import PeriodicBoundary

    def make_scheme(grid, bc):
        rec = make_reconstruction_from_name(rec_name)
        return advection.make_scheme_from_name("godunov", rec=rec, velocity=torch.full_like(grid.x, 1.0))

    eoc = ps.EOCRecorder(name=f"advection_godunov_{rec_name}")
    for n in resolutions:
        dt = 8.0 * (2.0 / n) ** (5.0 / 3.0)
        h, err = _evolve(make_scheme, lambda grid: PeriodicBoundary(),
                         lambda grid, t, x: funcs.ic_sine_sine(grid, x - 1.0 * t), n, dt=dt, tfinal=1.0)
        eoc.add_data_point(h, err)
    print(eoc)
    assert eoc.estimated_order >= order - 0.5


@pytest.mark.parametrize(("sname", "resolutions"), [
    ("rusanov", list(range(64, 128 + 1, 16))),
    ("lf", list(range(64, 128 + 1, 16))),
    ("eo", list(range(32, 128 + 1, 16))),
])
def test_burgers_convergence(sname: str, resolutions: list[int]) -> None:
    """tests/test_convergence.py:223-303: constant rec, Dirichlet Riemann data, alpha = 0.98, t = 1,
    EOC >= 0.9 (exercises the nu = df^(alpha-1) branch of the Rusanov / LF kernels)."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import burgers, funcs
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import make_dirichlet_boundary
    from pyshocks_b200.timestepping import predict_timestep_from_resolutions

    def make_scheme(grid, bc):
        return burgers.make_scheme_from_name(sname, rec=make_reconstruction_from_name("constant"), alpha=0.98)

    def make_bc(grid):
        return make_dirichlet_boundary(ga=lambda t, x: funcs.burgers_riemann(grid, t, x),
                                       gb=lambda t, x: funcs.burgers_riemann(grid, t, x))

    dt = predict_timestep_from_resolutions(-1.0, 1.0, resolutions, umax=10.0)
    eoc = ps.EOCRecorder(name=f"burgers_{sname}_constant")
    for n in resolutions:
        h, err = _evolve(make_scheme, make_bc, lambda grid, t, x: funcs.burgers_riemann(grid, t, x), n,
                         dt=dt, tfinal=1.0)
        eoc.add_data_point(h, err)
    print(eoc)
    assert eoc.estimated_order >= 1 - 0.1


@pytest.mark.parametrize(("name", "order", "resolutions"), [
    ("wenojs32", 3, list(range(192, 384 + 1, 32))),
    ("wenojs53", 5, list(range(32, 256 + 1, 32))),
    ("esweno32", 3, list(range(192, 384 + 1, 32))),  # tests/test_weno.py:293
])
def test_weno_smooth_reconstruction_order_cell_values(name: str, order: int, resolutions: list[int]) -> None:
    """tests/test_weno.py:288-338."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import BoundaryType
    from pyshocks_b200.reconstruction import make_reconstruction_from_name, reconstruct

    func = lambda x: torch.sin(2 * np.pi * x)  # noqa: E731
    eoc_l, eoc_r = ps.EOCRecorder(name="ul"), ps.EOCRecorder(name="ur")
    for n in resolutions:
        rec = make_reconstruction_from_name(name)
        grid = ps.make_uniform_cell_grid(-1.0, 1.0, n=n, nghosts=rec.stencil_width)
        quad = ps.make_leggauss_quadrature(grid, order=order + 1)
        u0 = ps.cell_average(quad, func)
        ref = func(grid.f)
        ul, ur = reconstruct(rec, grid, BoundaryType.Dirichlet, u0, u0, u0)
        eoc_l.add_data_point(grid.h, float(ps.rnorm(grid, ul, ref[:-1], p=float("inf"))))
        eoc_r.add_data_point(grid.h, float(ps.rnorm(grid, ur, ref[1:], p=float("inf"))))
    assert eoc_l.satisfied(order - 0.5)
    assert eoc_r.satisfied(order - 0.5)


@pytest.mark.parametrize(("cls_name", "order"), [("ForwardEuler", 1), ("SSPRK33", 3), ("RK44", 4), ("CKRK45", 4)])
def test_time_convergence(cls_name: str, order: int) -> None:
    """tests/test_timestepping.py:26-99: order >= expected - 0.1 on a scalar ODE (generic path)."""
    import pyshocks_b200 as ps
    import pyshocks_b200.timestepping as ts

    cls = getattr(ts, cls_name)
    eoc = ps.EOCRecorder(name=cls_name)
    tfinal = 4.0
    for n in range(2, 7):
        maxit, dt = ts.predict_maxit_from_timestep(tfinal, 1.0 / 2.0**n)
        stepper = cls(predict_timestep=lambda t, u, dt=dt: dt,
                      source=lambda t, u: torch.exp(-t) * torch.ones_like(u), checkpoint=None)
        u0 = torch.zeros(1, dtype=torch.float64, device="cuda")
        for event in ts.step(stepper, u0, maxit=maxit):
            pass
        exact = 1.0 - np.exp(-float(event.t))
        eoc.add_data_point(dt, abs(float(event.u[0]) - exact))
    assert eoc.estimated_order >= order - 0.1


def test_rk44_on_the_fused_rhs_converges() -> None:
    """RK44 / CKRK45 driving the fused apply_operator kernel (advection, WENOJS53, periodic)"""
    import pyshocks_b200 as ps
    import pyshocks_b200.timestepping as ts
    from pyshocks_b200 import advection, funcs
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import PeriodicBoundary

    errs = {}
    for name in ("RK44", "CKRK45", "SSPRK33"):
        grid = ps.make_uniform_cell_grid(a=-1.0, b=1.0, n=128, nghosts=3)
        bc = PeriodicBoundary()
        scheme = advection.Godunov(rec=make_reconstruction_from_name("wenojs53"), velocity=torch.ones_like(grid.x))
        quad = ps.make_leggauss_quadrature(grid, order=5)
        u0 = ps.cell_average(quad, lambda x: funcs.ic_sine_sine(grid, x))
        maxit, dt = ts.predict_maxit_from_timestep(0.25, 2.0e-3)
        stepper = getattr(ts, name)(predict_timestep=lambda t, u: dt, source=ps.bind_operator(scheme, grid, bc), checkpoint=None)
        for event in ts.step(stepper, u0, maxit=maxit):
            pass
        exact = ps.cell_average(quad, lambda x: funcs.ic_sine_sine(grid, x - 0.25))
        errs[name] = float(ps.rnorm(grid, event.u, exact, p=2, weighted=True))
    assert max(errs.values()) < 1e-5 and abs(errs["RK44"] - errs["SSPRK33"]) < 1e-6, errs


# }}}

# {{{ error behaviour of the boundary (SURVEY.md section 8b "errors")


def test_error_behaviour() -> None:
    import pyshocks_b200 as ps
    from pyshocks_b200 import burgers, timestepping
    from pyshocks_b200.checkpointing import InMemoryCheckpoint, load, save
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import PeriodicBoundary

    with pytest.raises(ValueError):
        ps.make_uniform_cell_grid(a=1.0, b=-1.0, n=16)  # grid.py:145-146
    with pytest.raises(ValueError):
        ps.make_uniform_cell_grid(a=-1.0, b=1.0, n=0)  # grid.py:148-149
    with pytest.raises(ValueError):
        make_reconstruction_from_name("muscl")  # outside the hot path
    with pytest.raises(ValueError):
        burgers.make_scheme_from_name("ssweno242")
    rec = make_reconstruction_from_name("wenojs53")
    grid = ps.make_uniform_cell_grid(a=-1.0, b=1.0, n=32, nghosts=3)
    bc = PeriodicBoundary()

    class Other(ps.SchemeBase):
        pass

    with pytest.raises(NotImplementedError):
        ps.apply_operator(Other(rec=rec), grid, bc, 0.0, grid.x)  # schemes.py:166
    with pytest.raises(NotImplementedError):
        ps.predict_timestep(Other(rec=rec), grid, bc, 0.0, grid.x)  # schemes.py:190
    scheme = burgers.Rusanov(rec=rec)
    with pytest.raises(TypeError):
        ps.apply_operator(scheme, grid, bc, 0.0, grid.x.cpu())  # no CPU fallback
    with pytest.raises(ValueError):
        ps.apply_operator(scheme, grid, bc, 0.0, grid.x[:-1].contiguous())  # scalar.py:227 shape assert
    small = ps.make_uniform_cell_grid(a=-1.0, b=1.0, n=32, nghosts=2)
    with pytest.raises(AssertionError):
        ps.apply_operator(scheme, small, bc, 0.0, small.x)  # reconstruction.py:369 nghosts >= stencil
    chk = InMemoryCheckpoint(basename="It")
    save(chk, 0, {"m": 0})
    with pytest.raises(KeyError):
        save(chk, 0, {"m": 0})  # checkpointing.py:123-124
    with pytest.raises(KeyError):
        load(chk, 3)  # checkpointing.py:134-135
    stepper = timestepping.SSPRK33(predict_timestep=lambda t, u: float("nan"), source=ps.bind_operator(scheme, grid, bc), checkpoint=None)
    with pytest.raises(ValueError):
        for _ in timestepping.step(stepper, torch.sin(grid.x), tfinal=float("inf")):  # timestepping.py:140-145
            pass
    with pytest.raises(ValueError):
        next(timestepping.adjoint_step(stepper, grid.x, maxit=1))  # timestepping.py:162-163


# }}}


# {{{ whole solve in one launch


@pytest.mark.parametrize("sname", ["rusanov", "lf", "godunov", "eo"])
def test_single_launch_solve_config1(sname: str) -> None:
    """examples/burgers.py config 1 through psk_solve_rows: STRICT reproduces the reference's dt
    history and final state bit for bit (rusanov, lf goldens); every scheme equals the
    step-by-step path bit for bit in STRICT mode and to 1e-12 in FAST mode."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import burgers, config, timestepping
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import PeriodicBoundary

    S = load_golden("solve_c1")
    rec = make_reconstruction_from_name("wenojs53")
    scheme = burgers.make_scheme_from_name(sname, rec=rec, alpha=1.0)
    grid = ps.make_uniform_cell_grid(a=-1.5, b=1.5, n=256, nghosts=3)
    bc = PeriodicBoundary()
    u0 = torch.from_numpy(S["rusanov_u0"]).cuda()
    i = grid.i_
    for math in ("strict", "fast"):
        config.set_math(math)
        res = timestepping.solve(scheme, grid, bc, u0, tfinal=1.0, theta=1.0, checkpoint=(math == "strict"))
        stepper = timestepping.SSPRK33(
            predict_timestep=lambda t_, u_: ps.predict_timestep(scheme, grid, bc, t_, u_),
            source=ps.bind_operator(scheme, grid, bc), checkpoint=None)
        dts = []
        for event in timestepping.step(stepper, u0, tfinal=1.0):
            dts.append(float(event.dt))
        assert int(res["iteration"][0]) == event.iteration
        if math == "strict":
            assert np.array_equal(host(res["dt"][0]), np.array(dts[1:]))
            assert torch.equal(res["u"], event.u)
            assert res["states"].shape[0] == event.iteration + 1
            assert torch.equal(res["states"][0][0], u0) and torch.equal(res["states"][-1][0], res["u"])
            if f"{sname}_uf" in S:
                assert np.array_equal(host(res["dt"][0]), S[f"{sname}_dt"][1:])
                assert np.array_equal(host(res["u"])[i], S[f"{sname}_uf"][i])
        else:
            assert max_rel(host(res["u"])[i], host(event.u)[i]) < 1e-12


def test_single_launch_solve_batched_rows_and_fixed_dt() -> None:
    from pyshocks_b200.ensemble import EnsembleSolver
    from pyshocks_b200.path import HotPath

    B, n, g, nsteps = 9, 700, 3, 15
    rng = np.random.default_rng(2)
    xh = (np.arange(n + 2 * g) - g + 0.5) / n
    u0 = torch.from_numpy(np.stack([
        rng.uniform(-0.5, 0.5) + np.sin(2 * np.pi * xh + rng.uniform(0, 6)) + 0.3 * np.sin(6 * np.pi * xh + rng.uniform(0, 6))
        for _ in range(B)])).cuda()
    for math in ("strict", "fast"):
        kw = dict(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=3.0 / n, eps=1e-12, math=math)
        ref = EnsembleSolver(batch=B, **kw).solve_fixed_dt(u0, 1e-4, nsteps).u
        hp = HotPath(**kw)
        u = u0.clone()
        out = hp.solve_rows(u, fixed_dt=1e-4, max_steps=nsteps)
        assert int(out["steps"].min()) == nsteps
        if math == "strict":
            assert torch.equal(u[:, g : g + n], ref[:, g : g + n])
        else:
            assert max_rel(host(u[:, g : g + n]), host(ref[:, g : g + n])) < 1e-12
    big = HotPath(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=8192, g=3, dx=1e-3, eps=1e-12)
    from pyshocks_b200._lib import PskError

    with pytest.raises(PskError):  # 5 * 8198 doubles do not fit in 227 KB of shared memory
        big.solve_rows(torch.zeros(8198, dtype=torch.float64, device="cuda"), fixed_dt=1e-4, max_steps=1)


# }}}


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("bc_name", ["periodic", "dirichlet"])
def test_burgers_esweno32_scheme_through_the_api(bc_name: str, math: str) -> None:
    """burgers/schemes.py:205-256 behind the generic functions: ``bind`` takes eps / delta from the
    grid (weno.py:263-281), ``numerical_flux`` / ``apply_operator`` / ``predict_timestep`` / ``advance``
    reproduce the golden vectors recorded from the reference (rough state: shock + exact zeros)."""
    from functools import partial

    import cases as C
    from common import load_golden, max_rel

    import pyshocks_b200 as ps
    import pyshocks_b200.timestepping as ts
    from pyshocks_b200 import burgers, config
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import PeriodicBoundary, make_dirichlet_boundary

    RHS, ADV = load_golden("rhs"), load_golden("advance")
    case = C.Case("burgers", "esweno32", "esweno32", bc_name, state="rough")
    k = case.key
    config.set_math(math)
    try:
        grid = ps.make_uniform_cell_grid(a=case.a, b=case.b, n=case.n, nghosts=case.g)
        with pytest.raises(TypeError):
            burgers.make_scheme_from_name("esweno32", rec=make_reconstruction_from_name("wenojs32"))
        scheme = burgers.make_scheme_from_name("esweno32", rec=make_reconstruction_from_name("esweno32"))
        if bc_name == "periodic":
            bc = PeriodicBoundary()
        else:
            bc = make_dirichlet_boundary(
                ga=lambda t, x: torch.from_numpy(C.dirichlet_values(case, float(t), x.cpu().numpy())).to(x.device))
        scheme = ps.bind(scheme, grid, bc)
        assert scheme.rec.eps == float(RHS[f"{k}_eps"]) and scheme.rec.delta == float(RHS[f"{k}_delta"])
        u = torch.from_numpy(RHS[f"{k}_u"]).cuda()
        w = ps.apply_boundary(bc, grid, case.t, u)
        F = ps.numerical_flux(scheme, grid, bc, case.t, w).cpu().numpy()
        L = ps.apply_operator(scheme, grid, bc, case.t, u).cpu().numpy()
        dt_cfl = float(ps.predict_timestep(scheme, grid, bc, case.t, u))
        assert dt_cfl == float(RHS[f"{k}_dt"])
        dt = float(ADV[f"{k}_dt"])
        stepper = ts.SSPRK33(predict_timestep=lambda t_, u_: dt, source=partial(ps.apply_operator, scheme, grid, bc),
                             checkpoint=None)
        out = ts.advance(stepper, dt, case.t, u).cpu().numpy()
        core = slice(3, -3)
        if math == "strict":
            assert np.array_equal(F[core], RHS[f"{k}_f"][core])
            assert np.array_equal(L[core], RHS[f"{k}_L"][core])
            assert np.array_equal(out[core], ADV[f"{k}_out"][core])
        else:
            assert max_rel(F, RHS[f"{k}_f"]) < 5e-13
            assert max_rel(L, RHS[f"{k}_L"]) < 5e-13
            assert max_rel(out, ADV[f"{k}_out"]) < 1e-13
        # the transposed ESWENO32 kernels (round 2), through the bound path, against reverse-mode differentiation of
        # the reference arithmetic (oracle/torch_twin.py)
        from common import oracle_setup
        from oracle import torch_twin as tt
        from pyshocks_b200.binding import hotpath_for

        hp = hotpath_for(scheme, grid, bc, case.t)
        v = np.random.default_rng(3).standard_normal(u.shape[0])
        got = hp.apply_operator_vjp(u, torch.from_numpy(v).cuda()).cpu().numpy()
        oscheme, ogrid, obc = oracle_setup(case)
        ref = tt.rhs_vjp(oscheme, ogrid, obc, case.t, u.cpu().numpy(), v)
        assert max_rel(got, ref) < 1e-10
    finally:
        config.set_math("fast")


def test_mixed_kind_two_sided_boundary_like_the_reference() -> None:
    """Dirichlet on the left, Neumann on the right: apply_boundary applies one side after the other
    (scalar.py:375-382) and equals the two one-kind fills on their sides; a scheme asking for
    bc.boundary_type raises NotImplementedError("Different boundaries on each side.") (scalar.py:369-370)."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import burgers
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import (DirichletBoundary, NeumannBoundary, TwoSidedBoundary, make_dirichlet_boundary,
                                      make_neumann_boundary)

    grid = ps.make_uniform_cell_grid(a=-1.0, b=1.0, n=40, nghosts=3)
    g = grid.nghosts
    u = torch.sin(3.0 * grid.x) + 0.3
    ga = lambda t, x: 0.25 + 0.0 * x + t  # noqa: E731
    gb = lambda t: 0.7  # noqa: E731
    mixed = TwoSidedBoundary(left=DirichletBoundary(side=-1, g=ga), right=NeumannBoundary(side=+1, g=gb))
    w = ps.apply_boundary(mixed, grid, 0.5, u)
    wd = ps.apply_boundary(make_dirichlet_boundary(ga), grid, 0.5, u)
    wn = ps.apply_boundary(make_neumann_boundary(gb), grid, 0.5, u)
    assert torch.equal(w[:g], wd[:g]) and torch.equal(w[g:-g], u[g:-g]) and torch.equal(w[-g:], wn[-g:])
    with pytest.raises(NotImplementedError, match="Different boundaries"):
        mixed.boundary_type
    scheme = burgers.Rusanov(rec=make_reconstruction_from_name("wenojs53"))
    with pytest.raises(NotImplementedError):
        ps.apply_operator(scheme, grid, mixed, 0.5, u)


@pytest.mark.parametrize("math", ["strict", "fast"])
def test_other_steppers_match_reference_advance(math: str) -> None:
    """ForwardEuler / RK44 / CKRK45 ``advance`` (timestepping.py:289-405) through the package API against the
    vectors recorded from the reference (tests/golden/steppers.npz): every RHS of a stage is one fused
    apply_operator launch, the combines are the reference's own array expressions, so STRICT results are
    bit-identical on every cell away from the zero-padded array ends."""
    from functools import partial

    import cases as C
    from common import load_golden, max_rel

    import pyshocks_b200 as ps
    import pyshocks_b200.timestepping as ts
    from pyshocks_b200 import advection, burgers, config, continuity
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import PeriodicBoundary, make_dirichlet_boundary

    G = load_golden("steppers")
    config.set_math(math)
    try:
        for case in C.stepper_cases():
            k = case.key
            grid = ps.make_uniform_cell_grid(a=case.a, b=case.b, n=case.n, nghosts=case.g)
            rec = make_reconstruction_from_name(case.rec)
            if case.equation == "burgers":
                scheme = burgers.make_scheme_from_name(case.flux, rec=rec, alpha=case.alpha)
            else:
                mod = advection if case.equation == "advection" else continuity
                scheme = mod.make_scheme_from_name(case.flux, rec=rec, velocity=torch.from_numpy(C.velocity_for(case)).cuda())
            if case.bc == "periodic":
                bc = PeriodicBoundary()
            else:
                bc = make_dirichlet_boundary(
                    ga=lambda t, x, case=case: torch.from_numpy(C.dirichlet_values(case, float(t), x.cpu().numpy())).to(x.device))
            scheme = ps.bind(scheme, grid, bc)
            u = torch.from_numpy(G[f"{k}_u"]).cuda()
            dt = float(G[f"{k}_dt"])
            for name in C.STEPPERS:
                stepper = getattr(ts, name)(predict_timestep=lambda t_, u_: dt,
                                            source=partial(ps.apply_operator, scheme, grid, bc), checkpoint=None)
                out = ts.advance(stepper, dt, case.t, u).cpu().numpy()
                ref = G[f"{k}_{name}"]
                core = slice(case.g + 2, -(case.g + 2))
                if math == "strict":
                    assert np.array_equal(out[core], ref[core]), (k, name)
                    assert max_rel(out, ref) < 1e-13, (k, name)
                else:
                    assert max_rel(out, ref) < 2e-13, (k, name)
    finally:
        config.set_math("fast")
