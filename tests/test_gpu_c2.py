"""BASELINE config 2 at its full size (drivers/advection-adjoint.py:219-328 with -s godunov -r wenojs53 -n 4096:
linear advection, Dirichlet exact-solution boundary, theta = 0.75, t = 1 -> 2731 steps), and the same path at
the sizes of the small golden sweeps:

* forward in ONE launch (timestepping.solve -> psk_solve_rows_tables: per-step dt and boundary-data tables)
  against the state the unmodified reference reaches (tests/golden/solve_c2_4096.npz);
* the reverse sweep in ONE call (timestepping.adjoint_solve -> psk_ssprk33_adjoint_sweep) against the
  step-by-step adjoint_step of the package (same kernels: bit-identical), against the reference's own
  adjoint_step vectors at N = 48, and -- at N = 4096 -- against reverse-mode differentiation of the reference
  arithmetic over the last steps of the sweep (the whole sweep would not fit an autograd graph)."""

from __future__ import annotations

from dataclasses import replace
from functools import partial

import numpy as np
import pytest
import torch

from common import load_golden, max_rel
from oracle import pyshocks_oracle as po
from oracle import torch_twin as tt

pytestmark = pytest.mark.gpu


def _driver(n: int, math: str = "fast"):
    import pyshocks_b200 as ps
    from pyshocks_b200 import advection, config, funcs
    from pyshocks_b200.reconstruction import make_reconstruction_from_name
    from pyshocks_b200.scalar import make_dirichlet_boundary

    config.set_math(math)
    rec = make_reconstruction_from_name("wenojs53")
    scheme = advection.make_scheme_from_name("godunov", rec=rec, velocity=None)
    grid = ps.make_uniform_cell_grid(a=-1.0, b=1.0, n=n, nghosts=scheme.stencil_width)
    quad = ps.make_leggauss_quadrature(grid, order=int(max(scheme.order, 1.0)) + 1)
    velocity = ps.cell_average(quad, partial(funcs.ic_constant, grid, c=1.0))
    scheme = replace(scheme, velocity=velocity)
    func_ic = partial(funcs.ic_sine, grid, k=1)
    u0 = ps.cell_average(quad, func_ic)
    bc = make_dirichlet_boundary(lambda t, x: func_ic(x - 1.0 * t))
    pbc = make_dirichlet_boundary(lambda t, x: torch.zeros_like(x))
    return scheme, grid, bc, pbc, u0


@pytest.mark.parametrize("math", ["strict", "fast"])
def test_config2_forward_in_one_launch_matches_the_reference(math: str) -> None:
    from pyshocks_b200 import config, timestepping

    S = load_golden("solve_c2_4096")
    try:
        scheme, grid, bc, pbc, u0 = _driver(4096, math)
        assert max_rel(u0.cpu().numpy(), S["u0"]) < 1e-15
        u0 = torch.from_numpy(S["u0"]).cuda()  # (device sin differs from libm in the last bit)
        res = timestepping.solve(scheme, grid, bc, u0, tfinal=1.0, theta=0.75, checkpoint=True)
        dts = res["dt"].reshape(-1).cpu().numpy()
        assert int(res["iteration"][0]) == 2731 and np.array_equal(dts, S["dt"][1:])  # the reference's dt history, bit for bit
        i = slice(3, 3 + 4096)
        for m, key in ((1, "u0001"), (100, "u0100"), (1000, "u1000"), (2731, "uf")):
            got = res["states"][m, 0, : 4102].cpu().numpy()
            # boundary data come from the device sin(): not the reference's bits, so no bitwise claim here
            assert max_rel(got[i], S[key][i]) < 1e-12, (m, max_rel(got[i], S[key][i]))
        assert torch.equal(res["u"], res["states"][2731, 0, :4102])
    finally:
        config.set_math("fast")


@pytest.mark.parametrize("n,tfinal", [(48, 0.5), (128, 1.0)])
def test_single_call_sweeps_equal_the_step_by_step_api(n: int, tfinal: float) -> None:
    """solve + adjoint_solve against step + adjoint_step: same kernels on the same data, same bits; and at
    N = 48 against the reference's own adjoint_step vectors (tests/golden/adjoint.npz)."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import timestepping
    from pyshocks_b200.checkpointing import InMemoryCheckpoint

    scheme, grid, bc, pbc, u0 = _driver(n)
    if n == 48:
        A = load_golden("adjoint")
        key = "advection_godunov_wenojs53_dirichlet"
        u0 = torch.from_numpy(A[f"{key}_u0"]).cuda()
    stepper = timestepping.SSPRK33(
        predict_timestep=ps.jit(lambda t_, u_: 0.75 * ps.predict_timestep(scheme, grid, bc, t_, u_)),
        source=ps.jit(lambda t_, u_: ps.apply_operator(scheme, grid, bc, t_, u_)),
        checkpoint=InMemoryCheckpoint(basename="Iteration"),
    )
    for event in timestepping.step(stepper, u0, tfinal=tfinal):
        pass
    uf, maxit = event.u, event.iteration
    ref_p = [ev.p for ev in timestepping.adjoint_step(
        stepper, uf, maxit=maxit, apply_boundary=lambda t, u, p: ps.apply_boundary(pbc, grid, t, p))]
    fwd = timestepping.solve(scheme, grid, bc, u0, tfinal=tfinal, theta=0.75, checkpoint=True)
    assert int(fwd["iteration"][0]) == maxit
    # the one-launch forward keeps the state in shared memory and evaluates whole rows per thread block: same
    # arithmetic per cell as the stage kernels
    i = slice(3, 3 + n)  # (the whole-step path leaves the ghost cells of its states zero)
    assert max_rel(fwd["u"].cpu().numpy()[i], uf.cpu().numpy()[i]) < 1e-13
    # reverse sweep from the SAME tape as the step-by-step path
    tape = torch.stack([stepper.checkpoint.storage[("Iteration", m)]["u"] for m in range(maxit + 1)])[:, None, :].contiguous()
    fwd2 = dict(fwd, states=tape)
    out = timestepping.adjoint_solve(scheme, grid, bc, fwd2, uf, p_boundary=pbc, history=True)
    assert torch.equal(out["p"], ref_p[-1])
    assert torch.equal(out["history"].flip(0), torch.stack(ref_p[1:]))  # history[m] = p after step m
    if n == 48:
        assert max_rel(torch.stack([ref_p[0]] + list(out["history"].flip(0))).cpu().numpy(), A[f"{key}_p"]) < 1e-12


def test_config2_adjoint_at_full_size() -> None:
    """N = 4096, 2731 steps: the single-call reverse sweep equals the step-by-step adjoint path (1e-13 over 100
    steps; bit for bit on identical array layouts, see the test above); its first 40 reverse steps equal torch
    autograd through the reference arithmetic (oracle/torch_twin.py) of the last 40 forward steps."""
    import pyshocks_b200 as ps
    from pyshocks_b200 import timestepping

    S = load_golden("solve_c2_4096")
    scheme, grid, bc, pbc, _ = _driver(4096)
    u0 = torch.from_numpy(S["u0"]).cuda()
    fwd = timestepping.solve(scheme, grid, bc, u0, tfinal=1.0, theta=0.75, checkpoint=True)
    nsteps = int(fwd["iteration"][0])
    uf = fwd["u"]
    out = timestepping.adjoint_solve(scheme, grid, bc, fwd, uf, p_boundary=pbc, history=True)
    hist = out["history"]  # hist[m] = p after reverse step m
    assert hist.shape == (nsteps, 4102) and bool(torch.isfinite(hist).all())
    assert torch.equal(out["p"], hist[0])
    # step-by-step API on the same tape, a slice of the sweep (the whole loop is ~2 s of Python)
    from pyshocks_b200.binding import ghost_data, hotpath_for

    hp = hotpath_for(scheme, grid, bc)
    fdts = [float(x) for x in fwd["dt"].reshape(-1).cpu().numpy()]
    ts = [0.0]
    for dt in fdts:
        ts.append(ts[-1] + dt)
    dts = [ts[m + 1] - ts[m] for m in range(nsteps)]  # adjoint_step: dt = t - chk["t"] (timestepping.py:200-202)
    p = ps.apply_boundary(pbc, grid, 0.0, uf)
    for m in range(nsteps - 1, nsteps - 101, -1):
        t, dt = ts[m], float(dts[m])
        ghosts = [ghost_data(bc, grid, tt) for tt in (t, t + dt, t + 0.5 * dt)]
        p = hp.ssprk33_step_adjoint(fwd["states"][m, 0, :4102], torch.tensor([dt], dtype=torch.float64, device="cuda"), p,
                                    ghosts=ghosts)
        p = ps.apply_boundary(pbc, grid, t, p)
        # (the tape of the whole-step path has aligned rows: the sweep runs the warp kernels, this loop -- on plain
        # unaligned arrays -- the tile kernels: same derivative, another summation order)
        assert max_rel(p.cpu().numpy(), hist[m].cpu().numpy()) < 1e-13, m
    # autograd twin over the last k steps: p_{nsteps-k} = (d u_nsteps / d u_{nsteps-k})^T-chain with the BC on p
    k = 40
    ogrid = po.make_grid(-1.0, 1.0, 4096, 3)
    vel = po.cell_average(ogrid, lambda x: np.full_like(x, 1.0), 4)
    oscheme = po.Scheme("advection", "godunov", po.make_reconstruction("wenojs53"), velocity=vel)
    obc = po.Dirichlet(ga=lambda t, x: po.ic_sine(ogrid, x - 1.0 * t, k=1))
    pt = uf.cpu().numpy().copy()
    pt[:3] = 0.0
    pt[-3:] = 0.0
    for m in range(nsteps - 1, nsteps - 1 - k, -1):
        um = fwd["states"][m, 0, :4102].cpu().numpy()
        pt = tt.step_vjp(oscheme, ogrid, obc, float(dts[m]), ts[m], um, pt)
        pt[:3] = 0.0
        pt[-3:] = 0.0
    got = hist[nsteps - k].cpu().numpy()
    assert max_rel(got, pt) < 1e-12, max_rel(got, pt)
