"""Profiling target: a few reverse steps of BASELINE config 5 (4096 x 8192) in the scheme of the reference's own
burgers-adjoint driver (global Lax-Friedrichs, alpha = 0.995, Dirichlet rows): adjoint_lean_kernel<4, 3, LF, NU>."""
import sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver
batch, n, g = 4096, 8192, 3
h = 3.0 / n
xc = -1.5 + h * (np.arange(n + 2 * g) - g + 0.5)
s = EnsembleSolver(equation="burgers", flux="lf", rec="wenojs53", bc="dirichlet", n=n, g=g, dx=h, eps=1e-12, batch=batch,
                   nu=np.diff(xc) ** (0.995 - 1.0))
s.hp.set_ghost(np.full(2 * g, 0.5))
x = torch.linspace(0, 1, s.nx, device="cuda", dtype=torch.float64)
u0 = 0.5 + torch.sin(2 * np.pi * x + 0.37)[None, :] * (0.2 + 0.8 * torch.rand(batch, 1, device="cuda", dtype=torch.float64))
adj = AdjointEnsemble(s, nsteps=4, dt=0.4 * h / 1.5, segment=2)
adj.gradient_half_l2(u0)
torch.cuda.synchronize()
