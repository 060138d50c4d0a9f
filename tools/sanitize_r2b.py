"""Small run of the kernels added late in round 2 for compute-sanitizer (memcheck / racecheck): the whole-step kernel
on Neumann rows and with the viscosity of every face, the Lax-Friedrichs cluster kernel (1 to 8 CTAs per row, idle
warps, Dirichlet rows, alpha != 1, inactive rows), psk_rhs_axpby through the fused RK44 / CKRK45 stages, the row-end
pass of advance(SSPRK33), and the transposed ESWENO32 kernels."""
import sys
from functools import partial

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
import pyshocks_b200 as ps  # noqa: E402
import pyshocks_b200.timestepping as ts  # noqa: E402
from pyshocks_b200 import burgers  # noqa: E402
from pyshocks_b200.ensemble import EnsembleSolver  # noqa: E402
from pyshocks_b200.path import HotPath  # noqa: E402
from pyshocks_b200.reconstruction import make_reconstruction_from_name  # noqa: E402
from pyshocks_b200.scalar import PeriodicBoundary  # noqa: E402

G = 3
for n in (16, 300, 1000, 2100, 4300, 8700):
    B = 3
    x = np.linspace(0, 6.28, n + 2 * G)
    u0 = torch.from_numpy(0.5 + np.sin(x)[None, :].repeat(B, 0) * np.array([[1.0], [0.5], [0.2]])).cuda()
    nu = np.diff(-1.37 + 3.1 * (np.arange(n + 2 * G) - G + 0.5) / n) ** (0.995 - 1.0)
    for flux, bc, kw in (("rusanov", "neumann", {}), ("rusanov", "dirichlet", {"nu": nu}), ("rusanov", "neumann", {"nu": nu}),
                         ("lf", "periodic", {}), ("lf", "dirichlet", {}), ("lf", "dirichlet", {"nu": nu})):
        s = EnsembleSolver(batch=B, equation="burgers", flux=flux, rec="wenojs53", bc=bc, n=n, g=G, dx=3.0 / n, eps=1e-12, **kw)
        if bc != "periodic":
            s.hp.set_ghost(np.full((B, 2 * G), 0.3 if bc == "dirichlet" else 1e-3))
        s.solve_fixed_dt(u0, 1e-4, 2)
        assert s._fused, (n, flux, bc)
        if flux == "lf" and n == 1000:  # rows that finish early are carried over by the whole cluster
            s.solve_adaptive(u0, theta=0.8, tfinal=2e-3, cfl_scale=0.5 * (3.0 / n), check_every=2)
# the other steppers with their combines fused into the RHS kernel, and advance(SSPRK33) with its row-end pass
grid = ps.make_uniform_cell_grid(a=-1.5, b=1.5, n=512, nghosts=G)
scheme = burgers.Rusanov(rec=make_reconstruction_from_name("wenojs53"), alpha=1.0)
bc = PeriodicBoundary()
u = 0.3 + torch.sin(2 * np.pi * grid.x.cuda() / 3.0)[None, :].repeat(2, 1)
for name in ("ForwardEuler", "SSPRK33", "RK44", "CKRK45"):
    st = getattr(ts, name)(predict_timestep=lambda t_, u_: 1e-3, source=partial(ps.apply_operator, scheme, grid, bc), checkpoint=None)
    v = ts.advance(st, 1e-3, 0.0, u)
    v = ts.advance(st, 1e-3, 1e-3, v)
    assert bool(torch.isfinite(v).all())
# transposed ESWENO32 kernels
for flux in ("esweno32", "godunov"):
    hp = HotPath(equation="burgers", flux=flux, rec="esweno32", bc="periodic", n=200, g=2, dx=3.0 / 200, eps=1e-3, delta=1e-3)
    w = torch.from_numpy(0.5 + np.sin(np.linspace(0, 6.28, 204))[None, :].repeat(2, 0)).cuda()
    out = hp.apply_operator_vjp(w, torch.ones_like(w))
    assert bool(torch.isfinite(out).all())
torch.cuda.synchronize()
print("sanitize r2b target done")
