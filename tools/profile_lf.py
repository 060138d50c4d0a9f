"""A few whole-step launches of the Lax-Friedrichs cluster kernel on a 16384 x 4096 ensemble (for ncu).
    python tools/profile_lf.py [periodic|dirichlet] [alpha]"""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import EnsembleSolver  # noqa: E402

bc = sys.argv[1] if len(sys.argv) > 1 else "periodic"
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
B, N, G = 16384, 4096, 3
x = (torch.arange(N + 2 * G, device="cuda", dtype=torch.float64) - G + 0.5) / N
coef = torch.from_numpy(np.random.default_rng(0).uniform(0.2, 1.0, size=(B, 1))).cuda()
u0 = 0.3 + coef * torch.sin(2 * np.pi * x)[None, :]
kw = {}
if alpha != 1.0:
    kw["nu"] = np.diff(-1.37 + 3.1 * (np.arange(N + 2 * G) - G + 0.5) / N) ** (alpha - 1.0)
s = EnsembleSolver(equation="burgers", flux="lf", rec="wenojs53", bc=bc, n=N, g=G, dx=3.0 / N, eps=1e-12, batch=B, **kw)
if bc != "periodic":
    s.hp.set_ghost(np.full(2 * G, 0.3))
s.load(u0)
dt = torch.full((1,), 0.4 * (3.0 / N) / 1.6, dtype=torch.float64, device="cuda")
s.solve_fixed_dt(None, dt, 4)
torch.cuda.synchronize()
assert s._fused
