"""Edge cases of the kernels' launch geometry and layout handling: tiny rows, rows shorter than a
warp chunk, ghost layers wider than the stencil, odd strides / unaligned views, batches beyond the
65535 grid.y limit, and argument errors.  The checker is the C oracle (bit-exact in STRICT mode)."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle.c_oracle import COracle
from common import max_rel

pytestmark = pytest.mark.gpu


def _rows(batch: int, nx: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, nx)
    return np.stack([0.3 * rng.standard_normal() + np.sin(2 * np.pi * (x + rng.uniform())) for _ in range(batch)])


@pytest.mark.parametrize("n", [3, 4, 5, 7, 8, 33, 119, 120, 121, 127, 128, 129, 241, 1000])
@pytest.mark.parametrize("math", ["strict", "fast"])
def test_row_lengths_around_the_chunk_size(n: int, math: str) -> None:
    """n = 3 is the smallest periodic grid with g = 3; 120 / 128 are the warp-chunk sizes"""
    from pyshocks_b200.ensemble import EnsembleSolver

    g, B, nsteps = 3, 4, 3
    dx = 1.0 / n
    u0 = _rows(B, n + 2 * g, n)
    dt = 0.2 * dx
    ref = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, batch=B,
                  dx=dx, eps=1e-12).solve_fixed_dt(u0, dt, nsteps)
    s = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=dx,
                       eps=1e-12, batch=B, math=math)
    out = s.solve_fixed_dt(torch.from_numpy(u0).cuda(), dt, nsteps).u.cpu().numpy()
    i = slice(g, g + n)
    if math == "strict":
        assert np.array_equal(out[:, i], ref[:, i])
    else:
        assert max_rel(out[:, i], ref[:, i]) < 1e-12


@pytest.mark.parametrize("g", [3, 4, 6])
def test_ghost_layer_wider_than_the_stencil(g: int) -> None:
    """grid.nghosts >= rec.stencil_width is all the reference asks for (reconstruction.py:369)"""
    from pyshocks_b200.path import HotPath

    n, B = 150, 2
    dx = 1.0 / n
    u0 = _rows(B, n + 2 * g, 7)
    co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, batch=B, dx=dx, eps=1e-12)
    ref_rhs = co.apply_operator(u0)
    ref_step = co.ssprk33_step(u0, 1e-3)
    hp = HotPath(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=dx, eps=1e-12, math="strict")
    ud = torch.from_numpy(u0).cuda()
    assert np.array_equal(hp.apply_operator(ud).cpu().numpy(), ref_rhs)  # all nx rows, ghost rows included
    out = hp.ssprk33_step(ud, torch.tensor([1e-3], dtype=torch.float64, device="cuda"), ghost_rows=True)
    assert np.array_equal(out.cpu().numpy(), ref_step)
    # adjoint with a wide ghost layer goes through the tile kernel; transpose identity against the RHS
    hpf = HotPath(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=dx, eps=1e-12)
    rng = np.random.default_rng(0)
    v = torch.from_numpy(rng.standard_normal(u0.shape)).cuda()
    d = torch.from_numpy(np.sin(np.linspace(0, 9, u0.shape[1]))[None, :].repeat(B, 0)).cuda()
    jtv = hpf.apply_operator_vjp(ud, v)
    h = 1e-6
    jd = (hpf.apply_operator(ud + h * d) - hpf.apply_operator(ud - h * d)) / (2 * h)
    lhs, rhs = (v * jd).sum(dim=1), (jtv * d).sum(dim=1)
    assert float(((lhs - rhs).abs() / lhs.abs().clamp_min(1.0)).max()) < 1e-6


@pytest.mark.parametrize("offset", [0, 1, 2, 5])
def test_unaligned_views_and_odd_strides(offset: int) -> None:
    """any base alignment / row stride works (scalar access path); the aligned fast path and the
    general path give the same bits"""
    from pyshocks_b200.ensemble import EnsembleSolver
    from pyshocks_b200.path import HotPath

    n, g, B = 500, 3, 3
    nx = n + 2 * g
    dx = 1.0 / n
    u0 = _rows(B, nx, 11)
    dt = torch.tensor([2e-4], dtype=torch.float64, device="cuda")
    aligned = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=dx,
                             eps=1e-12, batch=B)
    ref = aligned.solve_fixed_dt(torch.from_numpy(u0).cuda(), dt, 1).u
    hp = HotPath(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=dx, eps=1e-12)
    store = torch.zeros((B, nx + 7 + offset), dtype=torch.float64, device="cuda")
    view = store[:, offset : offset + nx]
    view.copy_(torch.from_numpy(u0))
    out = hp.ssprk33_step(view, dt)
    assert torch.equal(out[:, g : g + n], ref[:, g : g + n])
    # adjoint on the same odd layout (tile kernel) == adjoint on the aligned layout (warp kernel)
    p = torch.from_numpy(_rows(B, nx, 12)).cuda()
    pv = torch.zeros_like(store)[:, offset : offset + nx]
    pv.copy_(p)
    a = hp.ssprk33_step_adjoint(view, dt, pv)
    pa, ua = aligned.new_states(2)
    pa.copy_(p)
    ua.copy_(torch.from_numpy(u0))
    b = aligned.hp.ssprk33_step_adjoint(ua, dt, pa)
    assert max_rel(a.cpu().numpy(), b.cpu().numpy()) < 1e-13


@pytest.mark.parametrize("batch", [65535, 65536, 70001])
def test_batches_around_the_grid_limit(batch: int) -> None:
    """rows map to grid.y (max 65535): larger batches are sliced or take the general kernel"""
    from pyshocks_b200.ensemble import EnsembleSolver

    n, g = 64, 3
    dx = 1.0 / n
    s = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=dx,
                       eps=1e-12, batch=batch)
    base = torch.from_numpy(_rows(1, n + 2 * g, 3)).cuda()
    scale = 1.0 + (torch.arange(batch, device="cuda", dtype=torch.float64) % 7)[:, None] * 0.125
    s.solve_fixed_dt(base * scale, 1e-4, 2)
    one = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=dx,
                         eps=1e-12, batch=1)
    for r in (0, 6, 65534, batch - 1):
        one.solve_fixed_dt(base * scale[r : r + 1], 1e-4, 2)
        assert torch.equal(one.u[0, g : g + n], s.u[r, g : g + n])


def test_argument_errors() -> None:
    from pyshocks_b200._lib import PskError
    from pyshocks_b200.path import HotPath

    hp = HotPath(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=32, g=3, dx=0.1, eps=1e-12)
    u = torch.zeros(38, dtype=torch.float64, device="cuda")
    dt = torch.tensor([1e-3], dtype=torch.float64, device="cuda")
    with pytest.raises(PskError):
        hp.stage(1, u, u, u, dt)  # uout must not alias uin
    with pytest.raises(ValueError):
        hp.apply_operator(torch.zeros(40, dtype=torch.float64, device="cuda"))
    with pytest.raises(TypeError):
        hp.apply_operator(torch.zeros(38, dtype=torch.float32, device="cuda"))
    with pytest.raises(ValueError):
        hp.apply_operator(torch.zeros((2, 76), dtype=torch.float64, device="cuda")[:, ::2])  # non-unit stride
    with pytest.raises(PskError):  # periodic grid with fewer cells than ghosts
        HotPath(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=2, g=3, dx=0.1, eps=1e-12).apply_operator(
            torch.zeros(8, dtype=torch.float64, device="cuda"))
    with pytest.raises(PskError):  # Dirichlet without ghost data
        HotPath(equation="burgers", flux="rusanov", rec="wenojs53", bc="dirichlet", n=32, g=3, dx=0.1, eps=1e-12).apply_operator(u)
    with pytest.raises(ValueError):
        HotPath(equation="advection", flux="godunov", rec="wenojs53", bc="periodic", n=32, g=3, dx=0.1, eps=1e-12)
