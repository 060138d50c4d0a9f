"""Size-independent properties at (or near) the BASELINE.json sizes, where the CPU oracle is too
slow to be the checker: conservation, row independence, shift equivariance, mirror symmetry,
FAST-vs-STRICT agreement, decomposition invariance and the adjoint dot-product identity."""

from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = 3


def _solver(batch: int, n: int, math: str = "fast", flux: str = "rusanov"):
    from pyshocks_b200.ensemble import EnsembleSolver

    return EnsembleSolver(equation="burgers", flux=flux, rec="wenojs53", bc="periodic", n=n, g=G,
                          dx=3.0 / n, eps=1e-12, batch=batch, math=math)


def _ensemble_ic(batch: int, n: int, seed: int = 20261017) -> torch.Tensor:
    gen = torch.Generator(device="cuda").manual_seed(seed)
    c = torch.rand(batch, 1, generator=gen, device="cuda", dtype=torch.float64) - 0.5
    xh = ((torch.arange(n + 2 * G, device="cuda", dtype=torch.float64) - G + 0.5) / n)[None, :]
    u = c.repeat(1, n + 2 * G)
    for k in range(1, 5):
        a = torch.rand(batch, 1, generator=gen, device="cuda", dtype=torch.float64) / k
        ph = 2 * np.pi * torch.rand(batch, 1, generator=gen, device="cuda", dtype=torch.float64)
        u += a * torch.sin(2 * np.pi * k * xh + ph)
    return u


def test_config3_full_size_conservation_and_row_independence() -> None:
    """B = 65536 x N = 4096 (BASELINE configs[2]): sum_i u_i is conserved per row (flux-difference
    form, periodic), and randomly chosen rows equal their single-row solves bit for bit."""
    B, n, nsteps = 65536, 4096, 10
    s = _solver(B, n)
    u0 = _ensemble_ic(B, n)
    dt = 0.4 * (3.0 / n) / float(u0.abs().max())
    mass0 = u0[:, G : G + n].sum(dim=1)
    s.solve_fixed_dt(u0, dt, nsteps)
    mass1 = s.u[:, G : G + n].sum(dim=1)
    scale = u0[:, G : G + n].abs().sum(dim=1)
    assert float(((mass1 - mass0).abs() / scale).max()) < 1e-13
    assert bool(torch.isfinite(s.u[:, G : G + n]).all())
    one = _solver(1, n)
    for r in (0, 1, 31337, B - 1):
        one.solve_fixed_dt(u0[r : r + 1], dt, nsteps)
        assert torch.equal(one.u[0, G : G + n], s.u[r, G : G + n])


def test_shift_equivariance_is_bitwise() -> None:
    """periodic problem: rolling the initial data by k cells rolls the solution by k cells; the
    per-cell arithmetic does not depend on where a cell sits in a warp / chunk / row"""
    B, n, nsteps, k = 8, 4096, 15, 1237
    u0 = _ensemble_ic(B, n, seed=3)
    dt = 0.4 * (3.0 / n) / float(u0.abs().max())
    a = _solver(B, n)
    a.solve_fixed_dt(u0, dt, nsteps)
    u0s = u0.clone()
    u0s[:, G : G + n] = torch.roll(u0[:, G : G + n], k, dims=1)
    b = _solver(B, n)
    b.solve_fixed_dt(u0s, dt, nsteps)
    assert torch.equal(torch.roll(a.u[:, G : G + n], k, dims=1), b.u[:, G : G + n])


@pytest.mark.parametrize("flux", ["rusanov", "godunov", "eo", "lf"])
def test_mirror_symmetry(flux: str) -> None:
    """Burgers is invariant under u(x) -> -u(-x); the left/right-biased halves of the WENO code are
    written separately, so this checks one against the other"""
    B, n, nsteps = 4, 2048, 25
    u0 = _ensemble_ic(B, n, seed=9)
    dt = 0.3 * (3.0 / n) / float(u0.abs().max())
    a = _solver(B, n, flux=flux)
    a.solve_fixed_dt(u0, dt, nsteps)
    v0 = u0.clone()
    v0[:, G : G + n] = -torch.flip(u0[:, G : G + n], dims=(1,))
    b = _solver(B, n, flux=flux)
    b.solve_fixed_dt(v0, dt, nsteps)
    ref = -torch.flip(a.u[:, G : G + n], dims=(1,))
    err = float((b.u[:, G : G + n] - ref).abs().max() / ref.abs().max())
    assert err < 1e-12, err


def test_fast_vs_strict_large() -> None:
    """FAST (re-associated, FMA, fast reciprocal) against STRICT (the reference's operation order,
    bit-identical to the oracle) on 4096 rows x 4096 cells, 30 steps, shocks forming"""
    B, n, nsteps = 4096, 4096, 30
    u0 = _ensemble_ic(B, n, seed=11)
    dt = 0.4 * (3.0 / n) / float(u0.abs().max())
    f = _solver(B, n, "fast")
    f.solve_fixed_dt(u0, dt, nsteps)
    st = _solver(B, n, "strict")
    st.solve_fixed_dt(u0, dt, nsteps)
    ref = st.u[:, G : G + n]
    err = float((f.u[:, G : G + n] - ref).abs().max() / ref.abs().max())
    print(f"FAST vs STRICT after {nsteps} steps on {B} x {n}: max rel {err:.3e}")
    assert err < 1e-12


def test_adaptive_ensemble_matches_rowwise_reference_loop() -> None:
    """per-row adaptive dt (device-side step control, fused CFL maxima): every row follows the
    loop of timestepping.step on its own; rows finish after different numbers of steps"""
    B, n = 64, 512
    u0 = _ensemble_ic(B, n, seed=5)
    u0 *= torch.linspace(0.2, 3.0, B, device="cuda", dtype=torch.float64)[:, None]
    s = _solver(B, n, "strict")
    res = s.solve_adaptive(u0, theta=0.9, tfinal=0.05, cfl_scale=0.5 * (3.0 / n), check_every=4)
    assert len(set(res.steps_per_row.tolist())) > 4
    one = _solver(1, n, "strict")
    for r in (0, 17, B - 1):
        rr = one.solve_adaptive(u0[r : r + 1], theta=0.9, tfinal=0.05, cfl_scale=0.5 * (3.0 / n), check_every=1)
        assert rr.steps == int(res.steps_per_row[r])
        assert torch.equal(one.u[0, G : G + n], s.u[r, G : G + n])
        assert float(one.t[0]) == float(s.t[r])


def test_config4_like_single_grid_conservation() -> None:
    """one periodic grid of 2^27 cells (config 4 is 2^30; same kernels, same code path)"""
    n, nsteps = 1 << 27, 5
    s = _solver(1, n)
    x = (torch.arange(n, device="cuda", dtype=torch.float64) + 0.5) / n
    s.u[0, G : G + n] = 0.5 + torch.sin(2 * np.pi * x)
    del x
    m0 = float(s.u[0, G : G + n].sum())
    s.solve_fixed_dt(None, 0.4 * (3.0 / n) / 1.5, nsteps)
    m1 = float(s.u[0, G : G + n].sum())
    assert abs(m1 - m0) / (0.5 * n) < 1e-13


def test_config5_size_adjoint_dot_product_identity() -> None:
    """B = 4096 x N = 8192 (BASELINE configs[4]), 6 steps: <w, dF(u) v> by central differences of
    the forward solve equals <dF(u)^T w, v> from the reverse sweep, row by row"""
    from pyshocks_b200.ensemble import AdjointEnsemble

    B, n, nsteps = 4096, 8192, 6
    s = _solver(B, n)
    u0 = _ensemble_ic(B, n, seed=20261018)
    dt = 0.4 * (3.0 / n) / float(u0.abs().max())
    # smooth directions: with rough (random) ones the finite difference itself is the inaccurate
    # side (the WENO weights are strongly nonlinear at the grid scale; measured FD error 6e-2 at
    # h = 1e-6 for white-noise directions, falling with h, while both adjoint kernels agree)
    v = torch.zeros_like(u0)
    w = torch.zeros_like(u0)
    xh = ((torch.arange(n, device="cuda", dtype=torch.float64) + 0.5) / n)[None, :]
    ph = torch.linspace(0.0, 6.0, B, device="cuda", dtype=torch.float64)[:, None]
    v[:, G : G + n] = torch.cos(2 * np.pi * 3 * xh + ph)
    w[:, G : G + n] = torch.sin(2 * np.pi * 2 * xh + 1.0 - ph)
    adj = AdjointEnsemble(s, nsteps=nsteps, dt=dt, segment=3)
    adj.forward(u0)
    jtw = adj.backward(w).clone()
    rhs = (jtw[:, G : G + n] * v[:, G : G + n]).sum(dim=1)
    h = 1e-5
    s.solve_fixed_dt(u0 + h * v, dt, nsteps)
    up = s.u.clone()
    s.solve_fixed_dt(u0 - h * v, dt, nsteps)
    lhs = (w[:, G : G + n] * (up[:, G : G + n] - s.u[:, G : G + n]) / (2 * h)).sum(dim=1)
    rel = (lhs - rhs).abs() / (w[:, G : G + n].norm(dim=1) * v[:, G : G + n].norm(dim=1))
    print(f"adjoint dot-product identity at {B} x {n}: max |lhs - rhs| / (|w| |v|) = {float(rel.max()):.3e}")
    assert float(rel.max()) < 1e-8
