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
        assert _lib.lib().psk_set_stage_variant(7062) == 0


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


@pytest.mark.parametrize("code", [7060, 7062, 7082])
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


@pytest.mark.parametrize("kw", [dict(bc="dirichlet"), dict(flux="lf"), dict(math="strict"), dict(rec="wenojs32")])
def test_other_schemes_keep_the_stage_launches(kw: dict) -> None:
    batch, n = 2, 300
    u0 = _ic(batch, n, seed=2)
    s = _solver(batch, n, **kw)
    if kw.get("bc") == "dirichlet":
        s.hp.set_ghost(np.zeros(2 * G))
    s.solve_fixed_dt(u0, 1e-4, 2)
    assert s._fused is False and s.launches >= 6
    assert bool(torch.isfinite(s.u[:, G : G + n]).all())
