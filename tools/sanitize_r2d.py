"""Small run of the last kernels of round 2 for compute-sanitizer: the lean adjoint stage kernel with the global
Lax-Friedrichs speed / the viscosity of every face (arg-max record, speed cotangent atomics), the Lax-Friedrichs cluster
kernel with stage outputs (third stage skipped), and the adjoint sweep without the no-op boundary launch."""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver  # noqa: E402

G = 3
for n in (16, 300, 1000, 4300):
    B = 3
    x = np.linspace(0, 6.28, n + 2 * G)
    u0 = torch.from_numpy(0.5 + np.sin(x + 0.37)[None, :].repeat(B, 0) * np.array([[1.0], [0.5], [0.2]])).cuda()
    xc = -1.5 + (3.0 / n) * (np.arange(n + 2 * G) - G + 0.5)
    for flux, bc, alpha in (("lf", "dirichlet", 0.995), ("lf", "dirichlet", 1.0), ("lf", "periodic", 1.0), ("rusanov", "dirichlet", 0.995)):
        nu = None if alpha == 1.0 else np.diff(xc) ** (alpha - 1.0)
        s = EnsembleSolver(batch=B, equation="burgers", flux=flux, rec="wenojs53", bc=bc, n=n, g=G, dx=3.0 / n, eps=1e-12, nu=nu)
        if bc == "dirichlet":
            s.hp.set_ghost(np.full((B, 2 * G), 0.3))
        adj = AdjointEnsemble(s, nsteps=5, dt=1e-4, segment=2)
        J, grad = adj.gradient_half_l2(u0)
        assert bool(torch.isfinite(grad).all()), (n, flux, bc, alpha)
torch.cuda.synchronize()
print("sanitize r2d target done")
