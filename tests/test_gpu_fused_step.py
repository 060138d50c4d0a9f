"""The whole-step kernel (psk_ssprk33_step: the three SSPRK33 stages of timestepping.py:312-320 in one
launch, stage values in registers) against the three stage launches it replaces: same bits, for
fixed and adaptive time steps, graphs and the host-to-host call; and against the C oracle."""

from __future__ import annotations

import contextlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = 3


@contextlib.contextmanager
def whole_step(code: int):
    """7000 = three stage launches, 7000 + 10 R + shape = whole-step kernel (tuning switch of the ABI)"""
    from pyshocks_b200 import _lib

    assert _lib.lib().psk_set_stage_variant(code) == 0
    try:
        yield
    finally:
        assert _lib.lib().psk_set_stage_variant(7066) == 0  # the default shape


def _solver(batch: int, n: int, **kw):
    from pyshocks_b200.ensemble import EnsembleSolver

    args = dict(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=G, dx=3.0 / n, eps=1e-12,
                batch=batch, math="fast")
    args.update(kw)
    return EnsembleSolver(**args)


def _ic(batch: int, n: int, seed: int) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    x = (np.arange(n + 2 * G) - G + 0.5) / n
    u = np.stack([rng.uniform(-0.5, 0.5) + sum(rng.uniform(0, 1 / k) * np.sin(2 * np.pi * k * x + rng.uniform(0, 6.28))
                                               for k in range(1, 5)) for _ in range(batch)])
    u[:, G + n // 3 : G + n // 2] += 0.6
    return torch.from_numpy(u).cuda()


@pytest.mark.parametrize("code", [7060, 7061, 7062, 7064, 7066, 7068, 7082])
@pytest.mark.parametrize("batch,n,nsteps", [(5, 4096, 7), (3, 1000, 4), (2, 172, 3), (4, 50, 5), (1, 16, 2)])
def test_whole_step_equals_three_stage_launches(code: int, batch: int, n: int, nsteps: int) -> None:
    u0 = _ic(batch, n, seed=n + nsteps)
    dt = 0.4 * (3.0 / n) / float(u0.abs().max())
    with whole_step(7000):
        a = _solver(batch, n)
        a.solve_fixed_dt(u0, dt, nsteps)
        assert a._fused is False and a.launches == 3 * nsteps
    with whole_step(code):
        b = _solver(batch, n)
        b.solve_fixed_dt(u0, dt, nsteps)
        assert b._fused is True and b.launches == nsteps
    assert torch.equal(a.u[:, G : G + n], b.u[:, G : G + n])


@pytest.mark.parametrize("flux", ["godunov", "eo"])
@pytest.mark.parametrize("batch,n,nsteps", [(4, 2048, 5), (3, 333, 4)])
def test_whole_step_other_burgers_fluxes(flux: str, batch: int, n: int, nsteps: int) -> None:
    u0 = _ic(batch, n, seed=n)
    dt = 0.4 * (3.0 / n) / float(u0.abs().max())
    with whole_step(7000):
        a = _solver(batch, n, flux=flux)
        a.solve_fixed_dt(u0, dt, nsteps)
        assert a._fused is False
    b = _solver(batch, n, flux=flux)
    b.solve_fixed_dt(u0, dt, nsteps)
    assert b._fused is True and b.launches == nsteps
    assert torch.equal(a.u[:, G : G + n], b.u[:, G : G + n])


def test_whole_step_against_the_oracle() -> None:
    from oracle.c_oracle import COracle

    batch, n, nsteps = 6, 1500, 12
    u0 = _ic(batch, n, seed=1)
    dt = 0.4 * (3.0 / n) / float(u0.abs().max())
    s = _solver(batch, n)
    s.solve_fixed_dt(u0, dt, nsteps)
    assert s._fused is True
    co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=G, batch=batch, dx=3.0 / n,
                 eps=1e-12)
    ref = co.solve_fixed_dt(u0.cpu().numpy(), dt, nsteps)[:, G : G + n]
    got = s.u[:, G : G + n].cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()  # FAST tolerance


def test_adaptive_solve_is_unchanged() -> None:
    batch, n = 9, 600
    u0 = _ic(batch, n, seed=4)
    kw = dict(theta=0.8, tfinal=0.04, cfl_scale=0.5 * (3.0 / n), check_every=4, record_dt=True)
    with whole_step(7000):
        a = _solver(batch, n)
        ra = a.solve_adaptive(u0, **kw)
    b = _solver(batch, n)
    rb = b.solve_adaptive(u0, **kw)
    assert b._fused is True
    assert ra.steps == rb.steps and np.array_equal(ra.steps_per_row, rb.steps_per_row)
    assert np.array_equal(ra.dt_history, rb.dt_history)
    assert torch.equal(ra.t, rb.t)
    assert torch.equal(a.u[:, G : G + n], b.u[:, G : G + n])


@pytest.mark.parametrize("nsteps", [1, 6, 9])
def test_graph_replay_and_host_to_host_call(nsteps: int) -> None:
    batch, n = 12, 2048
    u0 = _ic(batch, n, seed=8)
    dt = 0.4 * (3.0 / n) / float(u0.abs().max())
    a = _solver(batch, n)
    a.solve_fixed_dt(u0, dt, nsteps)
    ref = a.u[:, G : G + n].clone()
    b = _solver(batch, n)
    b.solve_fixed_dt(u0, dt, nsteps, graph=True)
    assert torch.equal(b.u[:, G : G + n], ref)
    b.solve_fixed_dt(u0, dt, nsteps, graph=True)  # replay of the captured graph on fresh data
    assert torch.equal(b.u[:, G : G + n], ref)
    c = _solver(batch, n)
    host_in = u0.cpu().pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    c.solve_fixed_dt_host(host_in, host_out, dt, nsteps, groups=3, streams=2)
    torch.cuda.synchronize()
    assert torch.equal(host_out[:, G : G + n].cuda(), ref)
    assert torch.equal(c.u[:, G : G + n], ref)


@pytest.mark.parametrize("kw", [dict(math="strict"), dict(rec="wenojs32"), dict(flux="lf", bc="neumann")])
def test_other_schemes_keep_the_stage_launches(kw: dict) -> None:
    batch, n = 2, 300
    u0 = _ic(batch, n, seed=2)
    s = _solver(batch, n, **kw)
    if kw.get("bc") == "neumann":
        s.hp.set_ghost(np.zeros(2 * G))
    s.solve_fixed_dt(u0, 1e-4, 2)
    assert s._fused is False and s.launches >= 6
    assert bool(torch.isfinite(s.u[:, G : G + n]).all())


@pytest.mark.parametrize("equation,flux", [("burgers", "rusanov"), ("burgers", "godunov"), ("burgers", "eo"),
                                           ("advection", "godunov"), ("continuity", "godunov")])
@pytest.mark.parametrize("batch,n,nsteps,per_row", [(4, 4096, 5, True), (3, 333, 4, False), (2, 172, 3, True), (1, 16, 2, False)])
@pytest.mark.parametrize("bc", ["dirichlet", "neumann"])
def test_whole_step_on_rows_with_boundary_data(equation: str, flux: str, batch: int, n: int, nsteps: int, per_row: bool,
                                               bc: str) -> None:
    """psk_ssprk33_step_bc: Dirichlet / Neumann rows (Burgers fluxes, advection, continuity) in one launch per step,
    the same bits as three stage launches; EnsembleSolver picks it by itself (time-independent data of set_ghost)"""
    rng = np.random.default_rng(n)
    u0 = _ic(batch, n, seed=n + 1)
    kw = {}
    if equation != "burgers":
        x = (np.arange(n + 2 * G) - G + 0.5) / n
        kw["velocity"] = 1.0 + 0.4 * np.sin(2 * np.pi * x + 0.2)
    ghost = rng.uniform(-0.3, 0.3, size=(batch, 2 * G) if per_row else (2 * G,)) * (1.0 if bc == "dirichlet" else 3.0 / n)
    dt = 0.3 * (3.0 / n) / max(float(u0.abs().max()), 1.5)
    with whole_step(7000):
        a = _solver(batch, n, equation=equation, flux=flux, bc=bc, **kw)
        a.hp.set_ghost(ghost)
        a.solve_fixed_dt(u0, dt, nsteps)
        assert a._fused is False and a.launches == 3 * nsteps
    b = _solver(batch, n, equation=equation, flux=flux, bc=bc, **kw)
    b.hp.set_ghost(ghost)
    b.solve_fixed_dt(u0, dt, nsteps)
    assert b._fused is True and b.launches == nsteps
    assert torch.equal(a.u[:, G : G + n], b.u[:, G : G + n])


def test_whole_step_with_time_dependent_dirichlet_data() -> None:
    """the data of the three stage times differ (g(t, x) of the drivers): HotPath.step_fused(ghosts=...) against
    HotPath.ssprk33_step(ghosts=...) (three launches)"""
    from pyshocks_b200.path import HotPath

    n, batch = 1000, 3
    hp = HotPath(equation="burgers", flux="rusanov", rec="wenojs53", bc="dirichlet", n=n, g=G, dx=3.0 / n, eps=1e-12)
    s = _solver(batch, n, bc="dirichlet")
    u, out = s.new_states(2)
    u.copy_(_ic(batch, n, seed=5))
    rng = np.random.default_rng(0)
    ghosts = [rng.uniform(-0.3, 0.3, size=(batch, 2 * G)) for _ in range(3)]
    dt = torch.full((1,), 0.3 * (3.0 / n) / 1.5, dtype=torch.float64, device="cuda")
    assert hp.step_fused(u, out, dt, ghosts=ghosts)
    ref = hp.ssprk33_step(u, dt, ghosts=ghosts)
    assert torch.equal(out[:, G : G + n], ref[:, G : G + n])


def test_lax_friedrichs_without_the_reduction_pass() -> None:
    """global Lax-Friedrichs flux on periodic rows: every stage takes its speed from the fused maximum of the launch
    that produced its input (psk_ssprk33_stage_lf: 3 launches + 1 fill per step) -- the bits of the path with one
    reduction pass per stage (6 launches), also after a reload and from a CUDA graph"""
    batch, n, nsteps = 4, 2048, 6
    u0 = _ic(batch, n, seed=11)
    dt = 0.3 * (3.0 / n) / float(u0.abs().max())
    with whole_step(7000):  # (the stage launches: the whole-step cluster kernel is tested below)
        a = _solver(batch, n, flux="lf")
        a._lf_chain = False
        a.solve_fixed_dt(u0, dt, nsteps)
        assert a.launches == 6 * nsteps
        b = _solver(batch, n, flux="lf")
        b.solve_fixed_dt(u0, dt, nsteps)
        assert b._lf_chain and b.launches == 4 * nsteps + 1
        assert torch.equal(a.u[:, G : G + n], b.u[:, G : G + n])
        b.solve_fixed_dt(u0, dt, 2)       # reload: the chain starts again from a reduction of the new state
        b.solve_fixed_dt(None, dt, nsteps - 2)
        assert torch.equal(a.u[:, G : G + n], b.u[:, G : G + n])
        c = _solver(batch, n, flux="lf")
        c.solve_fixed_dt(u0, dt, nsteps, graph=True)
        assert torch.equal(a.u[:, G : G + n], c.u[:, G : G + n])


@pytest.mark.parametrize("bc", ["periodic", "dirichlet"])
@pytest.mark.parametrize("batch,n,nsteps", [(4, 4096, 5), (3, 2048, 4), (2, 5000, 3), (2, 10000, 3), (3, 16512, 2), (5, 300, 4),
                                            (2, 172, 3), (1, 16, 2)])
def test_whole_step_with_the_global_lax_friedrichs_flux(bc: str, batch: int, n: int, nsteps: int) -> None:
    """step_lf_cluster_kernel: one thread-block cluster per row, the row-wide max |w| of every stage input
    (scalar.py:277) exchanged through distributed shared memory behind a split cluster barrier -- one launch per
    step, the bits of the path with one reduction pass + one stage launch per stage (6 launches); rows of 1, 2, 4
    and 8 CTAs, with and without idle warps; Dirichlet rows take the ghost data of each stage into the maximum"""
    u0 = _ic(batch, n, seed=n + 3)
    dt = 0.3 * (3.0 / n) / max(float(u0.abs().max()), 1.0)
    rng = np.random.default_rng(n)
    # (Dirichlet data larger than the state: the speed then comes from the ghost cells)
    ghost = rng.uniform(-3.0, 3.0, size=(batch, 2 * G)) if bc == "dirichlet" else None
    with whole_step(7000):
        a = _solver(batch, n, flux="lf", bc=bc)
        a._lf_chain = False
        if ghost is not None:
            a.hp.set_ghost(ghost)
        a.solve_fixed_dt(u0, dt, nsteps)
        assert a._fused is False and a.launches == 6 * nsteps
    b = _solver(batch, n, flux="lf", bc=bc)
    if ghost is not None:
        b.hp.set_ghost(ghost)
    b.solve_fixed_dt(u0, dt, nsteps)
    assert b._fused is True and b.launches == nsteps
    assert torch.equal(a.u[:, G : G + n], b.u[:, G : G + n])
    c = _solver(batch, n, flux="lf", bc=bc)
    if ghost is not None:
        c.hp.set_ghost(ghost)
    c.solve_fixed_dt(u0, dt, nsteps, graph=True)
    assert torch.equal(a.u[:, G : G + n], c.u[:, G : G + n])


def test_lax_friedrichs_whole_step_adaptive_rows_and_long_rows() -> None:
    """rows that finish early are carried over by the whole cluster; the fused maximum of the new state feeds the
    next time step; rows beyond 8 x 12 windows fall back to the stage launches"""
    batch, n = 6, 1000
    u0 = _ic(batch, n, seed=21)
    kw = dict(theta=0.8, tfinal=0.02, cfl_scale=0.5 * (3.0 / n), check_every=4, record_dt=True)
    with whole_step(7000):
        a = _solver(batch, n, flux="lf")
        ra = a.solve_adaptive(u0, **kw)
    b = _solver(batch, n, flux="lf")
    rb = b.solve_adaptive(u0, **kw)
    assert b._fused is True
    assert ra.steps == rb.steps and np.array_equal(ra.steps_per_row, rb.steps_per_row)
    assert np.array_equal(ra.dt_history, rb.dt_history) and torch.equal(ra.t, rb.t)
    assert torch.equal(a.u[:, G : G + n], b.u[:, G : G + n])
    n = 16513
    c = _solver(1, n, flux="lf")
    c.solve_fixed_dt(_ic(1, n, seed=1), 1e-6, 1)
    assert c._fused is False


def test_lax_friedrichs_row_blocks_on_several_streams() -> None:
    """solve_fixed_dt_host cuts the batch into row blocks that run concurrently on several streams: every block
    needs its own speed buffer (the scratch is keyed by the launching stream); against the single-stream solve"""
    batch, n, nsteps = 64, 512, 5
    u0 = _ic(batch, n, seed=3)
    dt = 0.3 * (3.0 / n) / float(u0.abs().max())
    a = _solver(batch, n, flux="lf")
    a.solve_fixed_dt(u0, dt, nsteps)
    b = _solver(batch, n, flux="lf")
    host_in = u0.cpu().pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    b.solve_fixed_dt_host(host_in, host_out, dt, nsteps, groups=8, streams=4)
    torch.cuda.synchronize()
    assert torch.equal(host_out[:, G : G + n].cuda(), a.u[:, G : G + n])


@pytest.mark.parametrize("equation", ["advection", "continuity"])
def test_whole_step_for_linear_equations_on_periodic_rows(equation: str) -> None:
    """one launch per step when the velocity array's ghost cells are the periodic images of its interior (its
    reconstruction is then periodic too and the bits are those of three stage launches); three stage launches,
    chosen by the binding itself, for a velocity array that is not periodic in its ghost cells"""
    batch, n, nsteps = 4, 2048, 5
    u0 = _ic(batch, n, seed=9)
    x = (np.arange(n) + 0.5) / n
    vi = 0.2 + np.sin(2 * np.pi * x + 0.3)
    vel = np.concatenate([vi[n - G :], vi, vi[:G]])
    dt = 0.3 * (3.0 / n) / 1.3
    with whole_step(7000):
        a = _solver(batch, n, equation=equation, flux="godunov", velocity=vel)
        a.solve_fixed_dt(u0, dt, nsteps)
        assert a._fused is False and a.launches == 3 * nsteps
    b = _solver(batch, n, equation=equation, flux="godunov", velocity=vel)
    assert b.hp._vel_periodic
    b.solve_fixed_dt(u0, dt, nsteps)
    assert b._fused is True and b.launches == nsteps
    assert torch.equal(a.u[:, G : G + n], b.u[:, G : G + n])
    xg = (np.arange(n + 2 * G) - G + 0.5) / n
    c = _solver(batch, n, equation=equation, flux="godunov", velocity=1.0 + 0.3 * xg)  # not periodic
    assert not c.hp._vel_periodic
    c.solve_fixed_dt(u0, dt, nsteps)
    assert c._fused is False and c.launches == 3 * nsteps


@pytest.mark.parametrize("equation,flux,bc", [("burgers", "rusanov", "periodic"), ("burgers", "godunov", "dirichlet"),
                                              ("burgers", "eo", "periodic"), ("advection", "godunov", "periodic"),
                                              ("continuity", "godunov", "dirichlet"), ("burgers", "lf", "periodic"),
                                              ("burgers", "rusanov", "neumann")])
@pytest.mark.parametrize("batch,n", [(3, 1000), (1, 128), (2, 4096), (2, 100)])
def test_advance_keeps_the_ghost_cell_byproducts_of_the_reference(equation: str, flux: str, bc: str, batch: int, n: int) -> None:
    """HotPath.ssprk33_advance (what ``advance(SSPRK33)`` launches): the whole-step kernel for the interior plus the
    three stage launches on the 2 x 32 cells next to the row ends -- the FULL array, ghost cells included, carries
    the bits of three full stage launches with ghost_rows (the by-products the reference's advance leaves there,
    timestepping.py:312-320 over schemes.py:339-346); schemes without a whole-step kernel run the three launches"""
    from pyshocks_b200.path import HotPath

    kw = {}
    if equation != "burgers":
        vi = 0.2 + np.sin(2 * np.pi * (np.arange(n) + 0.5) / n + 0.3)
        kw["velocity"] = np.concatenate([vi[n - G :], vi, vi[:G]])
    hp = HotPath(equation=equation, flux=flux, rec="wenojs53", bc=bc, n=n, g=G, dx=3.0 / n, eps=1e-12, **kw)
    s = _solver(batch, n)
    (u,) = s.new_states(1)
    u.copy_(_ic(batch, n, seed=n + batch))
    rng = np.random.default_rng(1)
    ghosts = None if bc == "periodic" else [rng.uniform(-0.3, 0.3, size=2 * G) for _ in range(3)]
    dt = torch.full((1,), 0.3 * (3.0 / n) / 2.0, dtype=torch.float64, device="cuda")
    ref = hp.ssprk33_step(u, dt, ghosts=ghosts, ghost_rows=True)
    out = hp.ssprk33_advance(u, dt, ghosts=ghosts)
    assert torch.equal(out, ref)
    if bc != "periodic":  # the data of set_ghost at all three stage times
        hp.set_ghost(ghosts[0])
        assert torch.equal(hp.ssprk33_advance(u, dt), hp.ssprk33_step(u, dt, ghost_rows=True))
    one = hp.ssprk33_advance(u[0], dt, ghosts=ghosts)  # a single row, 1-D
    assert torch.equal(one, ref[0])


@pytest.mark.parametrize("flux", ["rusanov", "lf"])
@pytest.mark.parametrize("bc", ["dirichlet", "neumann", "periodic"])
@pytest.mark.parametrize("batch,n,nsteps", [(3, 4096, 4), (2, 1000, 3), (2, 172, 2)])
def test_whole_step_with_the_viscosity_of_every_face(flux: str, bc: str, batch: int, n: int, nsteps: int) -> None:
    """alpha != 1 (scalar.py:231-234; the burgers-adjoint driver of the reference runs Lax-Friedrichs with
    alpha = 0.995 on Dirichlet rows): nu = df ** (alpha - 1) per face of the array -- its entries differ in the last
    bit, diff of computed cell centres -- multiplies the speed of every face.  Rows with boundary data: one launch per
    step, the bits of the stage launches.  Periodic rows (the face at the seam has two entries) and Lax-Friedrichs on
    Neumann rows keep the stage launches."""
    x = -1.37 + 3.1 * (np.arange(n + 2 * G) - G + 0.5) / n
    nu = np.diff(x) ** (0.995 - 1.0)
    assert np.unique(nu).size > 1
    u0 = _ic(batch, n, seed=n + 7)
    rng = np.random.default_rng(n)
    ghost = None if bc == "periodic" else rng.uniform(-0.3, 0.3, size=(batch, 2 * G)) * (1.0 if bc == "dirichlet" else 3.0 / n)
    dt = 0.3 * (3.0 / n) / max(float(u0.abs().max()), 1.0) / float(nu.max())

    def solver():
        s = _solver(batch, n, flux=flux, bc=bc, nu=nu)
        if ghost is not None:
            s.hp.set_ghost(ghost)
        return s

    with whole_step(7000):
        a = solver()
        a._lf_chain = False
        a.solve_fixed_dt(u0, dt, nsteps)
        assert a._fused is False
    b = solver()
    b.solve_fixed_dt(u0, dt, nsteps)
    fused = bc == "dirichlet" or (bc == "neumann" and flux == "rusanov")
    assert b._fused is fused and (b.launches == nsteps) == fused
    assert torch.equal(a.u[:, G : G + n], b.u[:, G : G + n])
    one = _solver(batch, n, flux=flux, bc=bc)  # nu = 1 is a different scheme
    if ghost is not None:
        one.hp.set_ghost(ghost)
    one.solve_fixed_dt(u0, dt, nsteps)
    assert float((one.u[:, G : G + n] - b.u[:, G : G + n]).abs().max()) > 1e-9
