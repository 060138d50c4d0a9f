"""Ensemble forward + reverse sweep (BASELINE config 5 at parity size): the gradient of
J = 1/2 ||u(T)||^2 from the reverse sweep against a finite-difference directional derivative of
the GPU forward solve and against autograd through the torch twin of the oracle; the two-level
tape (segment thinning + recompute) against the full tape."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from common import max_rel
from oracle import pyshocks_oracle as po
from oracle import torch_twin as tt

pytestmark = pytest.mark.gpu


def _setup(batch: int, n: int, math: str = "fast"):
    from pyshocks_b200.ensemble import EnsembleSolver

    g = 3
    grid = po.make_grid(-1.5, 1.5, n, g)
    rng = np.random.default_rng(20261018)
    xh = (grid.x - grid.a) / (grid.b - grid.a)
    u0 = np.stack([
        rng.uniform(-0.5, 0.5) + sum(rng.uniform(0, 1 / k) * np.sin(2 * np.pi * k * xh + rng.uniform(0, 2 * np.pi)) for k in range(1, 5))
        for _ in range(batch)
    ])
    solver = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g,
                            dx=grid.h, eps=1e-12, batch=batch, math=math)
    dt = 0.4 * grid.h / np.abs(u0).max()
    return solver, grid, u0, dt


def test_gradient_vs_autograd_twin() -> None:
    from pyshocks_b200.ensemble import AdjointEnsemble

    batch, n, nsteps = 3, 64, 12
    solver, grid, u0, dt = _setup(batch, n)
    adj = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, segment=5)
    J, grad = adj.gradient_half_l2(torch.from_numpy(u0).cuda())
    J, grad = J.cpu().numpy(), grad.cpu().numpy()
    scheme = po.Scheme("burgers", "rusanov", po.make_reconstruction("wenojs53"))
    bc = po.Periodic()
    i = grid.interior
    for b in range(batch):
        u = torch.from_numpy(u0[b]).clone().requires_grad_(True)
        x = u
        for _ in range(nsteps):
            x = tt.ssprk33_advance(lambda t_, y: tt.apply_operator(scheme, grid, bc, t_, y), dt, 0.0, x)
        Jb = 0.5 * (x[i] ** 2).sum()
        (gb,) = torch.autograd.grad(Jb, u)
        assert abs(J[b] - float(Jb)) < 1e-12 * abs(float(Jb))
        # ghost cells of u0 are overwritten by the boundary condition: zero gradient there
        assert max_rel(grad[b], gb.numpy()) < 1e-12


def test_gradient_vs_finite_differences_of_gpu_forward() -> None:
    from pyshocks_b200.ensemble import AdjointEnsemble

    batch, n, nsteps = 4, 256, 40
    solver, grid, u0, dt = _setup(batch, n)
    adj = AdjointEnsemble(solver, nsteps=nsteps, dt=dt)
    J0, grad = adj.gradient_half_l2(torch.from_numpy(u0).cuda())
    grad = grad.cpu().numpy().copy()
    rng = np.random.default_rng(1)
    d = rng.standard_normal(u0.shape)
    d[:, : grid.g] = 0.0
    d[:, grid.nx - grid.g :] = 0.0
    h = 1e-6

    def J(u):
        solver.load(torch.from_numpy(u).cuda())
        uT = solver.solve_fixed_dt(None, dt, nsteps).u
        return (0.5 * (uT[:, grid.g : grid.g + n] ** 2).sum(dim=1)).cpu().numpy()

    fd = (J(u0 + h * d) - J(u0 - h * d)) / (2 * h)
    an = (grad * d).sum(axis=1)
    assert np.max(np.abs(fd - an) / np.maximum(np.abs(an), 1e-3)) < 1e-6, (fd, an)


@pytest.mark.parametrize("segment", [1, 4, 7, 25])
def test_two_level_tape_equals_full_tape(segment: int) -> None:
    from pyshocks_b200.ensemble import AdjointEnsemble

    batch, n, nsteps = 5, 128, 25
    solver, grid, u0, dt = _setup(batch, n)
    ref = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, segment=nsteps)
    _, g_ref = ref.gradient_half_l2(torch.from_numpy(u0).cuda())
    g_ref = g_ref.clone()
    solver2, *_ = _setup(batch, n)
    adj = AdjointEnsemble(solver2, nsteps=nsteps, dt=dt, segment=segment)
    _, g2 = adj.gradient_half_l2(torch.from_numpy(u0).cuda())
    assert torch.equal(g2, g_ref)  # recomputed states are bit-identical to the stored ones


@pytest.mark.parametrize("segment", [1, 3, 25])
def test_fused_recompute_gives_the_same_gradient(segment: int) -> None:
    """reverse sweep with (k1, k2[, next state]) recomputed in one launch: same bits as with stage launches"""
    from pyshocks_b200.ensemble import AdjointEnsemble

    batch, n, nsteps = 5, 128, 25
    solver, grid, u0, dt = _setup(batch, n)
    ref = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, segment=segment, fused_recompute=False)
    _, g_ref = ref.gradient_half_l2(torch.from_numpy(u0).cuda())
    g_ref = g_ref.clone()
    solver2, *_ = _setup(batch, n)
    adj = AdjointEnsemble(solver2, nsteps=nsteps, dt=dt, segment=segment, fused_recompute=True)
    _, g2 = adj.gradient_half_l2(torch.from_numpy(u0).cuda())
    assert adj.launches < ref.launches
    assert torch.equal(g2, g_ref)
