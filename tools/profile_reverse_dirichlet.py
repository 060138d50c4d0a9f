"""Profiling target: a few reverse steps of BASELINE config 5 (4096 x 8192) on DIRICHLET rows (reverse_step_kernel<20, 11, true>)."""
import sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver
batch, n = 4096, 8192
h = 3.0 / n
s = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="dirichlet", n=n, g=3, dx=h, eps=1e-12, batch=batch)
s.hp.set_ghost(np.full(6, 0.5))
x = torch.linspace(0, 1, s.nx, device="cuda", dtype=torch.float64)
u0 = 0.5 + torch.sin(2 * np.pi * x)[None, :] * torch.rand(batch, 1, device="cuda", dtype=torch.float64)
adj = AdjointEnsemble(s, nsteps=4, dt=0.4 * h / 1.5, segment=2)
adj.gradient_half_l2(u0)
torch.cuda.synchronize()
