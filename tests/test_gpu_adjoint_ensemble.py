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
    ref = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, segment=segment, fused_recompute=False, fused_reverse=False)
    _, g_ref = ref.gradient_half_l2(torch.from_numpy(u0).cuda())
    g_ref = g_ref.clone()
    solver2, *_ = _setup(batch, n)
    adj = AdjointEnsemble(solver2, nsteps=nsteps, dt=dt, segment=segment, fused_recompute=True, fused_reverse=False)
    _, g2 = adj.gradient_half_l2(torch.from_numpy(u0).cuda())
    assert adj.launches < ref.launches
    assert torch.equal(g2, g_ref)


@pytest.mark.parametrize("n,variant", [(256, 0), (250, 12), (1000, 16), (1000, 20), (4096, 0), (74, 12), (600, 24)])
def test_fused_reverse_step_matches_the_five_launch_path(n: int, variant: int) -> None:
    """psk_ssprk33_step_adjoint (one launch: recompute + 3 adjoint stages in shared memory) against
    psk_ssprk33_step_stages + 3 x psk_ssprk33_stage_adjoint on the same state: recomputed stage values
    bit-identical, cotangent equal to round-off (the transposed stencil is summed in a different order)."""
    from pyshocks_b200 import _lib as L
    from pyshocks_b200.ensemble import AdjointEnsemble

    batch = 5
    solver, grid, u0, dt = _setup(batch, n)
    assert L.lib().psk_set_reverse_variant(variant) == 0
    try:
        hp = solver.hp
        u, k1, k2, k1f, k2f, p, out, ref = solver.new_states(8)
        u.copy_(torch.from_numpy(u0).cuda())
        rng = np.random.default_rng(n)
        p[:, grid.g : grid.g + n] = torch.from_numpy(rng.standard_normal((batch, n))).cuda()
        dtt = torch.full((1,), dt, dtype=torch.float64, device="cuda")
        assert hp.reverse_step_fused(u, p, dtt, out, stages=(k1f, k2f))
        assert hp.step_fused_stages(u, k1, k2, None, dtt)
        i = slice(grid.g, grid.g + n)
        assert torch.equal(k1f[:, i], k1[:, i]) and torch.equal(k2f[:, i], k2[:, i])
        hp.ssprk33_step_adjoint(u, dtt, p, out=ref, stages=(k1, k2))
        assert max_rel(out[:, i].cpu().numpy(), ref[:, i].cpu().numpy()) < 2e-12  # both are ~1e-12 from the exact derivative
        assert float(out[:, : grid.g].abs().max()) == 0.0  # ghost cells are not written
    finally:
        L.lib().psk_set_reverse_variant(0)


@pytest.mark.parametrize("segment", [1, 2, 5])
def test_gradient_with_the_fused_reverse_step(segment: int) -> None:
    """whole sweeps: fused reverse step against the stage-by-stage sweep and against the autograd twin"""
    from pyshocks_b200.ensemble import AdjointEnsemble

    batch, n, nsteps = 3, 128, 11
    solver, grid, u0, dt = _setup(batch, n)
    ref = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, segment=segment, fused_reverse=False)
    _, g_ref = ref.gradient_half_l2(torch.from_numpy(u0).cuda())
    g_ref = g_ref.cpu().numpy().copy()
    solver2, *_ = _setup(batch, n)
    adj = AdjointEnsemble(solver2, nsteps=nsteps, dt=dt, segment=segment)
    assert adj.fused_reverse and adj.lam2 is None
    J, g2 = adj.gradient_half_l2(torch.from_numpy(u0).cuda())
    assert adj.launches < ref.launches
    assert max_rel(g2.cpu().numpy(), g_ref) < 2e-12
    scheme = po.Scheme("burgers", "rusanov", po.make_reconstruction("wenojs53"))
    i = grid.interior
    u = torch.from_numpy(u0[0]).clone().requires_grad_(True)
    x = u
    for _ in range(nsteps):
        x = tt.ssprk33_advance(lambda t_, y: tt.apply_operator(scheme, grid, po.Periodic(), t_, y), dt, 0.0, x)
    (gb,) = torch.autograd.grad(0.5 * (x[i] ** 2).sum(), u)
    assert max_rel(g2[0].cpu().numpy(), gb.numpy()) < 1e-12


def test_optimisation_loop_recovers_an_initial_condition() -> None:
    """Steepest descent on u0 with the discrete adjoint gradient (the loop the adjoint drivers are for,
    drivers/burgers-adjoint.py:269-315): the tracking objective J = 1/2 ||u(T; u0) - u(T; u0*)||^2 decreases
    monotonically from a perturbed start and the iterate moves towards u0*."""
    from pyshocks_b200.ensemble import AdjointEnsemble

    batch, n, nsteps = 4, 128, 30
    solver, grid, u0, dt = _setup(batch, n)
    adj = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, segment=4)
    ustar = torch.from_numpy(u0).cuda()
    target = adj.forward(ustar).clone()
    i = grid.interior
    xh = torch.from_numpy((grid.x - grid.a) / (grid.b - grid.a)).cuda()
    start = ustar + 0.05 * torch.sin(2 * np.pi * 2 * xh)[None, :]
    res = adj.optimize(start, niter=25, step=0.5, target=target)
    J = res.objective
    assert J.shape == (26, batch)
    assert (np.diff(J, axis=0) < 0).all(), J  # monotone descent for every row
    assert (J[-1] < 0.05 * J[0]).all(), (J[0], J[-1])
    e0 = (start[:, i] - ustar[:, i]).norm(dim=1)
    e1 = (res.u0[:, i] - ustar[:, i]).norm(dim=1)
    assert bool((e1 < e0).all())
    # target = None is the drivers' objective 1/2 ||u(T)||^2
    J0, g0 = adj.gradient(ustar)
    J1, g1 = adj.gradient_half_l2(ustar)
    assert torch.equal(J0, J1) and torch.equal(g0, g1)


@pytest.mark.parametrize("fused_reverse", [False, True])
def test_adjoint_ensemble_on_dirichlet_rows(fused_reverse: bool) -> None:
    """Dirichlet rows: the forward sweep runs psk_ssprk33_step_bc (one launch per step); the reverse sweep either
    recomputes k1, k2 with the same kernel (one launch) and applies three adjoint stage launches, or (the default)
    runs psk_ssprk33_step_adjoint_bc, one launch per reverse step; against autograd through the torch twin with
    the same (time-independent) boundary data"""
    from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver

    batch, n, nsteps, g = 3, 96, 8, 3
    grid = po.make_grid(-1.5, 1.5, n, g)
    rng = np.random.default_rng(5)
    xh = (grid.x - grid.a) / (grid.b - grid.a)
    u0 = np.stack([0.3 * b + np.sin(2 * np.pi * xh + b) for b in range(batch)])
    ghost = rng.uniform(-0.3, 0.3, size=2 * g)
    solver = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="dirichlet", n=n, g=g, dx=grid.h,
                            eps=1e-12, batch=batch)
    solver.hp.set_ghost(ghost)
    dt = 0.3 * grid.h / np.abs(u0).max()
    adj = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, segment=3, fused_reverse=None if fused_reverse else False)
    assert adj.fused_reverse == fused_reverse
    J, grad = adj.gradient_half_l2(torch.from_numpy(u0).cuda())
    assert adj._fused is True  # the forward sweep took the whole-step kernel
    if fused_reverse:  # forward nsteps + recomputed states inside the segments + one launch per reverse step
        assert adj.launches == nsteps + (2 + 2 + 1) + nsteps
    scheme = po.Scheme("burgers", "rusanov", po.make_reconstruction("wenojs53"))
    xg = np.concatenate([grid.x[:g], grid.x[-g:]])
    bc = po.Dirichlet(ga=lambda t, x: np.interp(x, xg, ghost))
    i = grid.interior
    for b in range(batch):
        u = torch.from_numpy(u0[b]).clone().requires_grad_(True)
        x = u
        for _ in range(nsteps):
            x = tt.ssprk33_advance(lambda t_, y: tt.apply_operator(scheme, grid, bc, t_, y), dt, 0.0, x)
        (gb,) = torch.autograd.grad(0.5 * (x[i] ** 2).sum(), u)
        assert max_rel(grad[b].cpu().numpy()[i], gb.numpy()[i]) < 1e-12


@pytest.mark.parametrize("n,variant", [(256, 0), (250, 12), (1000, 16), (1000, 20), (4096, 0), (74, 12), (600, 24), (8192, 0)])
def test_fused_reverse_step_on_dirichlet_rows_matches_the_five_launch_path(n: int, variant: int) -> None:
    """psk_ssprk33_step_adjoint_bc with per-row boundary data that differ at the three stage times (t, t + dt,
    t + dt / 2: timestepping.py:314-319) against psk_ssprk33_step_bc (stage outputs) + 3 x psk_ssprk33_stage_adjoint
    with the data of each stage: recomputed stage values bit-identical, cotangent equal to round-off, and the
    torch twin with the same boundary function for one row"""
    from pyshocks_b200 import _lib as L
    from pyshocks_b200.ensemble import EnsembleSolver

    batch, g = 4, 3
    grid = po.make_grid(-1.5, 1.5, n, g)
    rng = np.random.default_rng(n + 1)
    xh = (grid.x - grid.a) / (grid.b - grid.a)
    u0 = np.stack([0.2 * b + np.sin(2 * np.pi * xh + b) + 0.3 * np.cos(6 * np.pi * xh) for b in range(batch)])
    dt = 0.3 * grid.h / np.abs(u0).max()
    xg = np.concatenate([grid.x[:g], grid.x[-g:]])
    amp = rng.uniform(0.2, 0.9, size=batch)

    def data(b: int, t: float, x: np.ndarray) -> np.ndarray:
        return amp[b] * np.cos(2.0 * x - 40.0 * t) + 0.1 * b

    g3 = np.stack([np.stack([data(b, t, xg) for b in range(batch)]) for t in (0.0, dt, 0.5 * dt)])
    solver = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="dirichlet", n=n, g=g, dx=grid.h,
                            eps=1e-12, batch=batch)
    hp = solver.hp
    assert hp.reverse_step_supported()
    assert L.lib().psk_set_reverse_variant(variant) == 0
    try:
        u, k1, k2, k1f, k2f, p, out, lam2, lam1, ref = solver.new_states(10)
        u.copy_(torch.from_numpy(u0).cuda())
        i = slice(g, g + n)
        p[:, i] = torch.from_numpy(rng.standard_normal((batch, n))).cuda()
        dtt = torch.full((1,), dt, dtype=torch.float64, device="cuda")
        g3d = torch.from_numpy(g3).cuda()
        out.fill_(float("nan"))
        assert hp.reverse_step_fused(u, p, dtt, out, stages=(k1f, k2f), ghosts=g3d)
        # five launches: the stage values from the whole-step kernel, then the three transposed stages, each with
        # the boundary data of its own stage time
        k1r, k2r, _ = hp.ssprk33_step(u, dtt, ghosts=[g3[0], g3[1], g3[2]], keep_stages=True)
        assert torch.equal(k1f[:, i], k1r[:, i]) and torch.equal(k2f[:, i], k2r[:, i])
        hp.set_ghost(g3[2])
        hp.stage_adjoint(k2r, p, dtt, 2.0 / 3.0, lam2)
        hp.set_ghost(g3[1])
        hp.stage_adjoint(k1r, lam2, dtt, 1.0 / 4.0, lam1)
        hp.set_ghost(g3[0])
        hp.stage_adjoint(u, lam1, dtt, 1.0, ref, acc=p, c_acc=1.0 / 3.0, acc2=lam2, c_acc2=3.0 / 4.0)
        # both are ~1e-12 from the exact derivative (cells next to an extremum, beta ~ eps); the maximum over four
        # rows of 8192 cells came out at 2.5e-12
        assert max_rel(out[:, i].cpu().numpy(), ref[:, i].cpu().numpy()) < (2e-12 if n <= 4096 else 5e-12)
        assert bool(torch.isnan(out[:, :g]).all()) and bool(torch.isnan(out[:, g + n : g + n + g]).all())  # not written
        if n <= 1000:
            scheme = po.Scheme("burgers", "rusanov", po.make_reconstruction("wenojs53"))
            b = batch - 1
            bc = po.Dirichlet(lambda t, x, b=b: data(b, t, x))
            want = tt.step_vjp(scheme, grid, bc, dt, 0.0, u0[b], p[b, : n + 2 * g].cpu().numpy())
            assert max_rel(out[b, i].cpu().numpy(), want[i]) < 3e-12
    finally:
        L.lib().psk_set_reverse_variant(0)


@pytest.mark.parametrize("alpha", [1.0, 0.995])
@pytest.mark.parametrize("n", [96, 700, 4096])
def test_adjoint_ensemble_in_the_scheme_of_the_reference_driver(alpha: float, n: int) -> None:
    """The scheme of the reference's own burgers-adjoint driver (drivers/burgers-adjoint.py:68-97, :408: global
    Lax-Friedrichs flux, alpha = 0.995, Dirichlet rows): the forward sweep is one cluster launch per step, the reverse
    sweep recomputes k1, k2 with the same kernel (psk_ssprk33_step_bc with stage outputs: bit-identical to the stage
    launches) and applies three lean Lax-Friedrichs adjoint stages; against autograd through the torch twin."""
    from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver

    batch, nsteps, g = 3, 6, 3
    grid = po.make_grid(-1.5, 1.5, n, g)
    rng = np.random.default_rng(n)
    xh = (grid.x - grid.a) / (grid.b - grid.a)
    # (phases that keep the row's max |u| away from a tie between two cells: a tie decided by round-off would hand the
    # speed's cotangent to different cells here and in the twin)
    u0 = np.stack([0.3 * b + np.sin(2 * np.pi * xh + b + 0.37) for b in range(batch)])
    ghost = rng.uniform(-0.3, 0.3, size=2 * g)
    nu = None if alpha == 1.0 else np.diff(grid.x) ** (alpha - 1.0)
    dt = 0.3 * grid.h / np.abs(u0).max()
    grads = []
    for fused_recompute in (False, True):
        solver = EnsembleSolver(equation="burgers", flux="lf", rec="wenojs53", bc="dirichlet", n=n, g=g, dx=grid.h,
                                eps=1e-12, batch=batch, nu=nu)
        solver.hp.set_ghost(ghost)
        adj = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, segment=3, fused_recompute=fused_recompute)
        assert not adj.fused_reverse
        _, grad = adj.gradient_half_l2(torch.from_numpy(u0).cuda())
        assert adj._fused is True  # forward: one cluster launch per step
        grads.append((grad.clone(), adj.launches))
    # (the cotangent of a row's speed is summed with one atomic per warp: the order, hence the last bits, may differ)
    assert max_rel(grads[0][0].cpu().numpy(), grads[1][0].cpu().numpy()) < 1e-13
    assert grads[1][1] < grads[0][1]
    # the stage values of the one-launch recomputation are those of the stage launches, bit for bit
    hp = solver.hp
    u, k1, k2, k1s, k2s, un, uns = solver.new_states(7)
    u.copy_(torch.from_numpy(u0).cuda())
    dtt = torch.full((1,), dt, dtype=torch.float64, device="cuda")
    assert hp.step_fused_stages(u, k1, k2, un, dtt)
    hp.stage(1, u, u, k1s, dtt)
    hp.stage(2, u, k1s, k2s, dtt)
    hp.stage(3, u, k2s, uns, dtt)
    i = slice(g, g + n)
    assert torch.equal(k1[:, i], k1s[:, i]) and torch.equal(k2[:, i], k2s[:, i]) and torch.equal(un[:, i], uns[:, i])
    k1.zero_()
    k2.zero_()
    assert hp.step_fused_stages(u, k1, k2, None, dtt)  # third stage skipped
    assert torch.equal(k1[:, i], k1s[:, i]) and torch.equal(k2[:, i], k2s[:, i])
    if n <= 700:
        scheme = po.Scheme("burgers", "lf", po.make_reconstruction("wenojs53"), alpha=alpha)
        xg = np.concatenate([grid.x[:g], grid.x[-g:]])
        bc = po.Dirichlet(ga=lambda t, x: np.interp(x, xg, ghost))
        i = grid.interior
        for b in range(batch):
            u = torch.from_numpy(u0[b]).clone().requires_grad_(True)
            x = u
            for _ in range(nsteps):
                x = tt.ssprk33_advance(lambda t_, y: tt.apply_operator(scheme, grid, bc, t_, y), dt, 0.0, x)
            (gb,) = torch.autograd.grad(0.5 * (x[i] ** 2).sum(), u)
            # (random boundary data: the reconstruction next to the row ends is not smooth; 1.03e-12 measured)
            assert max_rel(grads[1][0][b].cpu().numpy()[i], gb.numpy()[i]) < 3e-12
