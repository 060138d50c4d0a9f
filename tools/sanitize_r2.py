"""Small run of every kernel added in round 2 for compute-sanitizer (memcheck / racecheck): the fused reverse
step (all run lengths, periodic rows and slabs), the whole-step kernel on Dirichlet rows / advection / continuity,
the whole step with the slab exchange inside the kernel, the table-driven solve and the adjoint sweep."""
import sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200 import _lib as L
from pyshocks_b200.distributed import PeerRing, PeerSlabAdjoint, PeerSlabSolver
from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver

G = 3
for n in (74, 300, 1000):
    B = 3
    kw = dict(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=G, dx=3.0 / n, eps=1e-12)
    x = np.linspace(0, 6.28, n + 2 * G)
    u0 = torch.from_numpy(0.5 + np.sin(x)[None, :].repeat(B, 0) * np.array([[1.0], [0.5], [0.2]])).cuda()
    for variant in (12, 16, 20, 24):
        L.lib().psk_set_reverse_variant(variant)
        adj = AdjointEnsemble(EnsembleSolver(batch=B, **kw), nsteps=3, dt=1e-3, segment=2)
        assert adj.fused_reverse
        adj.gradient_half_l2(u0)
        adj.optimize(u0, niter=1, step=0.1)
    L.lib().psk_set_reverse_variant(0)
    # whole-step kernel on Dirichlet rows, three equations
    vel = 1.0 + 0.3 * np.sin(x)
    for eq, flux in (("burgers", "rusanov"), ("burgers", "eo"), ("advection", "godunov"), ("continuity", "godunov")):
        s = EnsembleSolver(batch=B, equation=eq, flux=flux, rec="wenojs53", bc="dirichlet", n=n, g=G, dx=3.0 / n, eps=1e-12,
                           **({"velocity": vel} if eq != "burgers" else {}))
        s.hp.set_ghost(np.full((B, 2 * G), 0.3))
        s.solve_fixed_dt(u0, 1e-3, 2)
        assert s._fused
# slabs: whole step + exchange in one launch (plain and shifted chunk grid), slab adjoint
for n, world in ((6144, 2), (2 * (172 * 9 + 4), 2), (1536, 3)):
    ug = torch.from_numpy(0.5 + np.sin(np.linspace(0, 6.28, n))).cuda()
    dtt = torch.full((1,), 1e-4, dtype=torch.float64, device="cuda")
    slabs = [PeerSlabSolver(n_global=n, rank=r, world=world, dx=3.0 / n, whole_step=True, timeout_s=30.0) for r in range(world)]
    for r, sl in enumerate(slabs):
        sl.attach(PeerRing.local([t.mem for t in slabs], r))
    for sl in slabs:
        sl.load_interior(ug[sl.first : sl.first + sl.n_local])
    for _ in range(3):
        for sl in slabs:
            sl.step(dtt)
    for sl in slabs:
        assert sl.fused_step
        sl.join(); sl.check(); sl.ring = None; sl.solver = None; sl.mem.close()
    if n % (2 * world) == 0:
        ad = [PeerSlabAdjoint(n_global=n, rank=r, world=world, dx=3.0 / n, nsteps=3, timeout_s=30.0) for r in range(world)]
        for a in ad:
            a.attach_local(ad)
        for a in ad:
            a.forward_begin(ug[a.first : a.first + a.n_local])
        for m in range(3):
            for a in ad:
                a.forward_step(m, dtt)
        for a in ad:
            a.backward_begin(a.interior(3))
        for m in range(2, -1, -1):
            for a in ad:
                a.backward_step(m, dtt)
        for a in ad:
            a.check(); a.close()
# config-2 style: table-driven forward (both forms) and the adjoint sweep
from dataclasses import replace
from functools import partial
import pyshocks_b200 as ps
from pyshocks_b200 import advection, config, funcs, timestepping
from pyshocks_b200.reconstruction import make_reconstruction_from_name
from pyshocks_b200.scalar import make_dirichlet_boundary
for math in ("fast", "strict"):
    config.set_math(math)
    rec = make_reconstruction_from_name("wenojs53")
    scheme = advection.make_scheme_from_name("godunov", rec=rec, velocity=None)
    grid = ps.make_uniform_cell_grid(a=-1.0, b=1.0, n=200, nghosts=3)
    quad = ps.make_leggauss_quadrature(grid, order=4)
    scheme = replace(scheme, velocity=ps.cell_average(quad, partial(funcs.ic_constant, grid, c=1.0)))
    func_ic = partial(funcs.ic_sine, grid, k=1)
    u0 = ps.cell_average(quad, func_ic)
    bc = make_dirichlet_boundary(lambda t, x: func_ic(x - 1.0 * t))
    pbc = make_dirichlet_boundary(lambda t, x: torch.zeros_like(x))
    fwd = timestepping.solve(scheme, grid, bc, u0, tfinal=0.05, theta=0.75, checkpoint=True)
    timestepping.adjoint_solve(scheme, grid, bc, fwd, fwd["u"], p_boundary=pbc, history=True)
config.set_math("fast")
torch.cuda.synchronize()
print("sanitize r2 target done")
