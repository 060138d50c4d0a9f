"""A few whole-step launches on rows with boundary data, 16384 x 4096 (for ncu): python tools/profile_step_bc.py [dirichlet|neumann|periodic]"""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import EnsembleSolver  # noqa: E402

bc = sys.argv[1] if len(sys.argv) > 1 else "dirichlet"
B, N, G = 16384, 4096, 3
x = (torch.arange(N + 2 * G, device="cuda", dtype=torch.float64) - G + 0.5) / N
coef = torch.from_numpy(np.random.default_rng(0).uniform(0.2, 1.0, size=(B, 1))).cuda()
u0 = 0.3 + coef * torch.sin(2 * np.pi * x)[None, :]
s = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc=bc, n=N, g=G, dx=3.0 / N, eps=1e-12, batch=B)
if bc != "periodic":
    s.hp.set_ghost(np.full(2 * G, 0.3 if bc == "dirichlet" else 0.0))
s.load(u0)
dt = torch.full((1,), 0.4 * (3.0 / N) / 1.6, dtype=torch.float64, device="cuda")
s.solve_fixed_dt(None, dt, 4)
torch.cuda.synchronize()
assert s._fused
